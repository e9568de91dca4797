"""BASELINE configs[3]: Charades-STA-shaped grounding — 64 synthetic 30 s videos (60 frames @ 2 fps) with a 24-token query
encoded before frame 0 (test/inference.py:281-282), `stream_end_prob_threshold=1` => never generates: the output is 60
relevance + 60 informative scores per video.  Videos are round-robined over the ranks (parallel.videos_for_rank); each rank
decodes its videos as CONCURRENT streams in one decoder step (B videos x k frames per weight pass), the only exchange is
the gather of the scores to rank 0.  Run alone (1 GPU) or under torchrun.

  python tools/config4_grounding.py [n_videos=64] [streams_per_step=8] [frames_per_pass=10]"""
import json, os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmduet_b200.config import ModelConfig
from mmduet_b200.engine import DecoderEngine, VisionEngine
from mmduet_b200.parallel import gather_results, videos_for_rank
from mmduet_b200.random_init import random_state_dict, synthetic_frames

n_videos = int(sys.argv[1]) if len(sys.argv) > 1 else 64
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
N_FRAMES, PREFIX, QUERY = 60, 32, 24
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
cfg = ModelConfig()
sd = random_state_dict(cfg, seed=1234, device=dev, include_lm_head=False)
vis = VisionEngine(cfg, sd, dev)
tpf = vis.tokens_per_frame
ctx_len = PREFIX + QUERY + N_FRAMES * tpf
dec = DecoderEngine(cfg, sd, dev, n_pages=B * ((ctx_len + 63) // 64 + 1) + 2, max_tokens=B * (PREFIX + QUERY + tpf * k), max_context=ctx_len + 64)
mine = videos_for_rank(n_videos, world, rank)


_inputs = {}


def video_inputs(v):
    """uint8 frames resident in HBM + query ids (generated once, outside the timed region)."""
    if v not in _inputs:
        g = torch.Generator().manual_seed(1000 + v)
        _inputs[v] = (synthetic_frames(N_FRAMES, seed=100 + v, device=dev), torch.randint(0, 151643, (QUERY,), generator=g).tolist())
    return _inputs[v]


def run(videos, streams_per_step, frames_per_pass):
    """scores[v] = [60, 2] (informative, relevance) for every video of `videos`."""
    out = {}
    for b0 in range(0, len(videos), streams_per_step):
        group = videos[b0:b0 + streams_per_step]
        inputs = [video_inputs(v) for v in group]
        emb = vis.visual_embed(torch.cat([fr for fr, _ in inputs]), normalize=True).view(len(group), N_FRAMES * tpf, cfg.hidden)
        streams, L = [dec.new_stream() for _ in group], [0] * len(group)
        sc = [[] for _ in group]
        for f0 in range(0, N_FRAMES, frames_per_pass):
            nf = min(frames_per_pass, N_FRAMES - f0)
            items = []
            for i, (_, q) in enumerate(inputs):
                ids = (list(range(100, 100 + PREFIX)) + q) if f0 == 0 else []     # system prompt stand-in + the query turn
                items.append(dict(storage=streams[i], past=L[i], ids=ids, frames=emb[i, f0 * tpf:(f0 + nf) * tpf],
                                  score_rows=[len(ids) + tpf * (j + 1) - 1 for j in range(nf)]))
            o = dec.step(items, score="frame_ends")
            s = o["scores"].view(len(group), nf, 2)
            for i in range(len(group)):
                L[i] = o["views"][i].length
                sc[i].append(s[i])
        for i, v in enumerate(group):
            out[v] = torch.cat(sc[i], 0)
            streams[i].release()
    return out


for v in mine:
    video_inputs(v)
run(mine[:min(len(mine), B)], B, k)                       # warm-up
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
scores = run(mine, B, k)
torch.cuda.synchronize()
dt_local = time.perf_counter() - t0
if world > 1:
    dist.barrier()
dt = time.perf_counter() - t0
res = {v: s.cpu().tolist() for v, s in scores.items()}
allres = gather_results(res, dst=0) if world > 1 else [res]
if rank == 0:
    merged = {v: s for r in allres for v, s in r.items()}
    assert sorted(merged) == list(range(n_videos)) and all(len(s) == N_FRAMES for s in merged.values())
    # batching must not change a video's scores: video 0 alone, one frame per pass (the reference's schedule)
    solo = run([0], 1, 1)[0].cpu()
    diff = (solo - torch.tensor(merged[0])).abs().max().item()
    rep = {"config": "BASELINE configs[3] grounding", "world": world, "videos": n_videos, "frames": n_videos * N_FRAMES,
           "streams_per_step": B, "frames_per_pass": k, "seconds": round(dt, 3), "videos_per_s": round(n_videos / dt, 2),
           "frames_per_s": round(n_videos * N_FRAMES / dt, 1), "video0_batched_vs_solo_per_frame_maxdiff": diff,
           "video0_relevance_first5": [round(x[1], 4) for x in merged[0][:5]]}
    print(json.dumps(rep))
    assert diff < 2e-2, diff
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rep, open(f"gpurun_out/config4_w{world}.json", "w"))
if world > 1:
    dist.destroy_process_group()
