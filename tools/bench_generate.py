"""Row f1: greedy response generation (M = 1 decoder steps over the paged KV + last-row lm_head + device argmax with
repetition penalty) at a 3k-token context: ms per generated token against the weight-streaming floor
((13.05 GB decoder + 1.09 GB lm_head) / 6552 GB/s = 2.16 ms)."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmduet_b200.config import ModelConfig
from mmduet_b200.modeling_live import VideoHeadLiveLlavaQwenForCausalLM, fast_greedy_generate
from mmduet_b200.random_init import random_state_dict

dev = torch.device("cuda:0")
cfg = ModelConfig()
sd = random_state_dict(cfg, seed=1234, device=dev, include_lm_head=True)
model = VideoHeadLiveLlavaQwenForCausalLM(cfg, sd, device=dev, max_context=4096, max_step_tokens=512)
dec = model.decoder
g = torch.Generator(device=dev).manual_seed(3)
fill = (torch.randn(490, cfg.hidden, device=dev, generator=g) * 1.1).bfloat16()
st, L = dec.new_stream(), 0
for _ in range(6):
    out = dec.step([dict(storage=st, past=L, embeds=fill)], score="last")
    L = out["views"][0].length
view = out["views"][0]
res = {}
for n_new, pen in ((64, None), (64, 1.15)):
    ids = torch.zeros(1, n_new, dtype=torch.long, device=dev)
    x = model.get_input_embeddings()(torch.tensor([[151645, 198, 151644, 77091, 198]], device=dev))   # "<|im_end|>\n<|im_start|>assistant\n"
    for rep in range(2):                                  # first repetition warms up
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out_ids, v2, _ = fast_greedy_generate(model=model, inputs_embeds=x, past_key_values=view, eos_token_id=-1,
                                              inplace_output_ids=ids, repetition_penalty=pen, generated_token_ids=[])
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    res[f"penalty={pen}"] = {"tokens": int(out_ids.shape[1]), "ms_per_token": round(1e3 * dt / out_ids.shape[1], 3), "context": view.length}
res["weight_streaming_floor_ms"] = round((13.05e9 + 1.09e9) / 6552e9 * 1e3, 2)
print(json.dumps(res))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/bench_generate.json", "w"))
