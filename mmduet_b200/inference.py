"""Frame loop over the CUDA model behind the reference's class / method / attribute names:
  LiveInferForBenchmark    test/inference.py:20-313   (reset, set_fps, input_video_stream, input_query_stream,
                                                       _encode_frame, _encode_query, _generate_response, inference)
  LiveInferForDemo         demo/liveinfer.py:60-105   (encode_given_query, input_one_frame)

Structure (not the reference's): one `StreamSession` owns everything a video stream is — the queue of encoded frames, the
pending user queries, the KV view, the turn state that decides which template ids precede the next tokens, the video clock
and the `DecisionRule` (which scores count, which threshold, the running sum).  The session advances in *passes* of k >= 1
frames: one decoder launch sequence scores k frame ends, the rule is applied to them in order, and at the first frame that
asks for a response the speculative rest of the pass is undone (O(1) KV rollback, frames back to the queue).  k = 1 is the
reference's frame-by-frame loop; any k gives the same scores and decisions (tests/test_gpu_loop.py).  The two public
classes are thin facades that expose the session's state under the reference's attribute names.

Differences from the reference that do not change results: frame tokens stay on the device (the reference parks them on the
CPU and copies them back, test/inference.py:212,237); frame steps skip lm_head (its output is never read there); all scores
of a pass come back in ONE device->host copy instead of two .item() syncs per frame."""
import collections
import math
import threading
import time
from dataclasses import asdict

import torch

from .modeling_live import fast_greedy_generate

_EMPTY = ()


def template_ids(tokenizer, conversation, **flags):
    """tokenizer.apply_chat_template(..., return_tensors='pt') as a flat list of ids, whatever the tokenizer returns
    (transformers 4.x: a [1, n] tensor; 5.x: a BatchEncoding unless return_dict=False; SyntheticTokenizer: a tensor)."""
    try:
        out = tokenizer.apply_chat_template(conversation, return_tensors="pt", return_dict=False, **flags)
    except TypeError:
        out = tokenizer.apply_chat_template(conversation, return_tensors="pt", **flags)
    if hasattr(out, "keys"):
        out = out["input_ids"]
    return [int(i) for i in torch.as_tensor(out).reshape(-1).tolist()]


class DecisionRule:
    """When to speak (test/inference.py:289-301, demo/liveinfer.py:83-94): the sum of the selected heads' scores against a
    single-frame threshold (strict >) or a running-sum threshold that resets when crossed."""

    def __init__(self, score_heads, prob_threshold=None, sum_threshold=None, running_list_length=20):
        self.score_heads = list(score_heads)
        self.prob_threshold, self.sum_threshold = prob_threshold, sum_threshold
        self.running_list_length = running_list_length
        self.clear()

    def clear(self):
        self.recent = []      # the reference's stream_end_prob_list
        self.total = 0        # the reference's stream_end_score_sum

    def observe(self, scores):
        s = sum(v for k, v in scores.items() if k in self.score_heads)
        self.recent.append(s)
        self.total += s
        n = self.running_list_length
        if isinstance(n, int) and n > 0:
            del self.recent[:-n]
        speak = False
        if self.sum_threshold is not None and self.total > self.sum_threshold:
            speak, self.total = True, 0
        if self.prob_threshold is not None and s > self.prob_threshold:
            speak = True
        return speak


class StreamSession:
    """State and primitives of one video stream over the CUDA decoder."""

    def __init__(self, model, tokenizer, *, system_prompt, rule, remove_assistant_turns, repetition_penalty, eos_token_id,
                 max_new_tokens=200):
        self.model, self.tokenizer, self.rule = model, tokenizer, rule
        self.device = model.device
        self.hidden = model.config.hidden_size
        self.n_tok = model.vision.tokens_per_frame
        self.remove_assistant_turns = remove_assistant_turns
        self.repetition_penalty = repetition_penalty
        self.eos_token_id = eos_token_id
        self.out_ids = torch.zeros(1, max_new_tokens, device=self.device, dtype=torch.long)
        self.ids_system = template_ids(tokenizer, [{"role": "system", "content": system_prompt}])
        self.ids_stream = template_ids(tokenizer, [{}], add_stream_prompt=True)
        self.ids_generate = template_ids(tokenizer, [{}], add_stream_generation_prompt=True)
        self.frames_per_pass = 1
        self.fps = None
        self.lock = threading.Lock()   # the demo re-enters the step from another thread (demo/app.py:84-85)
        self.view = None
        self._copy_stream = None
        self.step_ms = None            # set to a list to record host wall-clock ms of every frame pass (bench.py)
        self.clear()

    # ---- state ----
    def clear(self):
        self.frames = collections.deque()       # (video time, [n_tok, hidden] device tensor)
        self.queries = collections.deque()      # (time, text)
        self.clock = 0                           # the reference's video_time (accumulated 1/fps increments)
        self.n_frames_seen = 0
        self.since_reply = 0
        self.role = "system"
        self.carry = _EMPTY                      # ids to feed before the next tokens (the reference's last_ids)
        if self.view:
            self.view.storage.release()          # pages go back to the pool
        self.view = None
        self.debug = []
        self.generated = []
        self.rule.clear()

    @property
    def context_len(self):
        return self.view.length if self.view else 0

    # ---- input ----
    def push_video(self, video_frames, batch=None):
        """uint8 [T,3,384,384] frames (what test/datasets.py yields; rescale/normalise happen inside the patch-embed kernel) or
        already-processed float pixel_values.  Frames are independent through the encoder, so the batch size (the reference
        uses 32, test/inference.py:208) does not change any value: the encoder's preferred batch is used, and host frames are
        uploaded batch by batch on a copy stream so that batch i+1 crosses PCIe while batch i is being encoded."""
        batch = batch or self.model.vision.MAX_BATCH
        n = len(video_frames)
        if video_frames.device == self.device:
            chunks = [(video_frames[b:b + batch], None) for b in range(0, n, batch)]
        else:
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=self.device)
            main = torch.cuda.current_stream(self.device)
            chunks = []
            with torch.cuda.stream(self._copy_stream):
                for b in range(0, n, batch):
                    dev = video_frames[b:b + batch].to(self.device, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self._copy_stream)
                    dev.record_stream(main)
                    chunks.append((dev, ev))
        for i, (dev, ev) in enumerate(chunks):
            if ev is not None:
                torch.cuda.current_stream(self.device).wait_event(ev)
            tokens = self.model.visual_embed(dev).split(self.n_tok)
            self.frames.extend(((i * batch + r) / self.fps, t) for r, t in enumerate(tokens))

    def push_queries(self, conversation):
        self.queries.extend((t["time"], t["content"]) for t in conversation if t["role"] == "user")

    # ---- decoder passes ----
    def _launch(self, ids, frames=None, score="none", score_rows=None, lm="none"):
        view = self.view if self.view else self.model.new_cache()
        item = dict(storage=view.storage, past=view.length, ids=list(ids), frames=frames)
        if score_rows is not None:
            item["score_rows"] = score_rows
        out = self.model.decoder.step([item], score=score, lm=lm)
        self.view = out["views"][0]
        return out

    def _frame_prefix(self):
        """Template ids in front of the next frame tokens: the system turn on an empty context, the stream prompt after an
        assistant turn that stays in the context, nothing while the stream continues."""
        if not self.view:
            return list(self.ids_system)
        if self.role == "assistant" and not self.remove_assistant_turns:
            return list(self.carry) + list(self.ids_stream)
        return []

    def _pass_len(self):
        k = min(self.frames_per_pass, len(self.frames))
        if self.queries and k > 1:
            # a query is encoded in front of the first frame whose time has reached it: stop the pass there
            t, t_query, n = self.clock, self.queries[0][0], 0
            while n < k and not (n > 0 and t >= t_query):
                n += 1
                t += 1 / self.fps
            k = n
        return max(k, 1)

    def frame_pass(self, k):
        """Appends the next k queued frames in ONE decoder pass.  Returns (frames, scores [k][2], context length after each)."""
        t0 = time.perf_counter() if self.step_ms is not None else 0.0
        taken = [self.frames.popleft() for _ in range(k)]
        prefix = self._frame_prefix()
        self.carry = prefix
        past, P, n = self.context_len, len(prefix), self.n_tok
        tokens = taken[0][1].view(-1, self.hidden) if k == 1 else torch.cat([f[1].view(-1, self.hidden) for f in taken], 0)
        out = self._launch(prefix, tokens, score="frame_ends", score_rows=[P + n * (j + 1) - 1 for j in range(k)])
        scores = out["scores"].tolist()                      # one device->host copy for the whole pass
        if self.step_ms is not None:
            self.step_ms.append((time.perf_counter() - t0) * 1e3)
        return taken, scores, [past + P + n * (j + 1) for j in range(k)]

    def undo_after(self, taken, lens, j):
        """Frames j+1.. of a pass were speculative: cut the context back to the end of frame j, re-queue them."""
        from .engine import CacheView
        self.view.storage.truncate(lens[j])
        self.view = CacheView(self.view.storage, lens[j])
        self.frames.extendleft(reversed(taken[j + 1:]))

    def saw_frame(self):
        self.n_frames_seen += 1
        self.since_reply += 1
        self.role = "stream"

    def query_turn(self, text):
        ids = template_ids(self.tokenizer, [{"role": "user", "content": text}], add_stream_query_prompt=self.role == "stream",
                           add_stream_prompt=True)
        out = self._launch(ids, lm="last")
        self.carry = [int(out["lm_logits"].argmax(dim=-1).item())]
        self.role = "user"

    def respond(self):
        """Greedy response after the generation prompt.  With remove_assistant_turns the cache returned by the generator is
        dropped, i.e. the next pass appends at the length the context had BEFORE the response (the transformers 4.44.2
        legacy-cache meaning of test/inference.py:265-269)."""
        before = self.view
        prompt = torch.tensor([self.ids_generate], device=self.device, dtype=torch.long)
        ids, after, self.generated = fast_greedy_generate(
            model=self.model, inputs_embeds=self.model.get_input_embeddings()(prompt), past_key_values=before,
            eos_token_id=self.eos_token_id, inplace_output_ids=self.out_ids, repetition_penalty=self.repetition_penalty,
            generated_token_ids=self.generated)
        if self.remove_assistant_turns:
            self.carry = _EMPTY
        else:
            self.view = after
            self.carry = [int(ids[0, -1])]
        self.since_reply = 0
        self.role = "assistant"
        return self.tokenizer.decode(ids[0], skip_special_tokens=True, clean_up_tokenization_spaces=True)

    # ---- the loop ----
    def advance(self, on_response):
        """One pass: due query, k frames, decisions in frame order.  on_response(time, text) is called for a response."""
        if self.queries and self.clock >= self.queries[0][0]:
            self.query_turn(self.queries.popleft()[1])
        k = self._pass_len()
        taken, scores, lens = self.frame_pass(k)
        for j in range(k):
            self.saw_frame()
            sc = {"informative_score": scores[j][0], "relevance_score": scores[j][1]}
            self.debug.append(dict(time=self.clock, **sc))
            speak = self.rule.observe(sc)
            if speak:
                if j + 1 < k:
                    self.undo_after(taken, lens, j)
                on_response(self.clock, self.respond())
            self.clock += 1 / self.fps
            if speak:
                break


def _forward_to_session(name, target):
    return property(lambda self: getattr(self.session, target), lambda self, v: setattr(self.session, target, v))


class LiveInferForBenchmark:
    def __init__(self, args, model=None, tokenizer=None) -> None:
        assert not (args.bf16 and args.fp16), "only one of --bf16 true and --fp16 true can be set"
        if not args.bf16:
            raise ValueError("the accelerated path computes in bf16 (run with --bf16 true, as every reference script does)")
        self.torch_dtype = torch.bfloat16
        if model is None:
            from . import build_model_and_tokenizer
            model, tokenizer = build_model_and_tokenizer(is_training=False, set_vision_inside=True, torch_dtype=self.torch_dtype, **asdict(args))
        self.model, self.tokenizer = model.eval(), tokenizer
        self.image_processor = model.get_vision_tower().image_processor
        self.device = model.device
        n_set = sum(v is not None for v in (args.threshold_z, args.stream_end_prob_threshold, args.stream_end_score_sum_threshold))
        if n_set != 1:
            raise ValueError('only one of --stream_end_prob_threshold, --threshold_z and --stream_end_score_sum_threshold can be set. '
                             f'However, they are: {args.stream_end_prob_threshold}, {args.threshold_z}, {args.stream_end_score_sum_threshold}')
        if args.threshold_z is not None:
            if args.first_n_frames_no_generate is None:
                raise ValueError('--first_n_frames_no_generate must be set when --threshold_z is set')
            raise NotImplementedError("--threshold_z is only implemented by the reference's DEPRECATED _call_for_streaming loop")
        eos = model.config.eos_token_id
        if eos is None:
            eos = getattr(tokenizer, "eos_token_id", None)
        rule = DecisionRule(args.score_heads.split(','), args.stream_end_prob_threshold, args.stream_end_score_sum_threshold,
                            args.running_list_length)
        self.session = StreamSession(model, tokenizer, system_prompt=args.system_prompt, rule=rule,
                                     remove_assistant_turns=args.remove_assistant_turns, repetition_penalty=args.repetition_penalty,
                                     eos_token_id=eos)
        # names the reference's scripts read
        self.hidden_size = model.config.hidden_size
        self.frame_resolution = model.config.frame_resolution
        self.frame_num_tokens = self.session.n_tok
        self.frame_v_placeholder = model.config.v_placeholder * self.frame_num_tokens
        self.system_prompt = args.system_prompt
        self.response_min_interval_frames = args.response_min_interval_frames
        self.threshold_z = args.threshold_z
        self.first_n_frames_no_generate = args.first_n_frames_no_generate
        self.consecutive_n_frames_threshold = args.consecutive_n_frames_threshold
        self.consecutive_n_frames = 0
        self.video_tensor = None
        if args.frame_fps > 0:
            self.set_fps(args.frame_fps)

    # the reference's attribute names, backed by the session
    frame_embeds_queue = _forward_to_session("frame_embeds_queue", "frames")
    query_queue = _forward_to_session("query_queue", "queries")
    video_time = _forward_to_session("video_time", "clock")
    frame_idx = _forward_to_session("frame_idx", "n_frames_seen")
    num_frames_no_reply = _forward_to_session("num_frames_no_reply", "since_reply")
    last_role = _forward_to_session("last_role", "role")
    past_key_values = _forward_to_session("past_key_values", "view")
    debug_data_list = _forward_to_session("debug_data_list", "debug")
    generated_token_ids = _forward_to_session("generated_token_ids", "generated")
    inplace_output_ids = _forward_to_session("inplace_output_ids", "out_ids")
    remove_assistant_turns = _forward_to_session("remove_assistant_turns", "remove_assistant_turns")
    repetition_penalty = _forward_to_session("repetition_penalty", "repetition_penalty")
    eos_token_id = _forward_to_session("eos_token_id", "eos_token_id")
    frames_per_step = _forward_to_session("frames_per_step", "frames_per_pass")
    _lock = _forward_to_session("_lock", "lock")

    def _ids_property(target):   # template ids as the [1, n] device tensors the reference keeps
        def get(self):
            return torch.tensor([list(getattr(self.session, target))], device=self.device, dtype=torch.long)

        def put(self, v):
            setattr(self.session, target, [int(i) for i in torch.as_tensor(v).reshape(-1).tolist()])
        return property(get, put)

    _start_ids = _ids_property("ids_system")
    _added_stream_prompt_ids = _ids_property("ids_stream")
    _added_stream_generation_ids = _ids_property("ids_generate")
    last_ids = _ids_property("carry")
    del _ids_property

    @property
    def score_heads(self):
        return self.session.rule.score_heads

    @property
    def stream_end_prob_threshold(self):
        return self.session.rule.prob_threshold

    @property
    def stream_end_score_sum_threshold(self):
        return self.session.rule.sum_threshold

    @property
    def running_list_length(self):
        return self.session.rule.running_list_length

    @property
    def stream_end_prob_list(self):
        return self.session.rule.recent

    @property
    def stream_end_score_sum(self):
        return self.session.rule.total

    @property
    def frame_fps(self):
        return self.session.fps

    @property
    def frame_interval(self):
        return 1 / self.session.fps

    def set_fps(self, fps=None, frame_interval=None):
        assert (fps is None) != (frame_interval is None), "give exactly one of fps / frame_interval"
        self.session.fps = fps if fps is not None else 1 / frame_interval

    def reset(self):
        self.session.clear()
        self.consecutive_n_frames = 0
        self.video_tensor = None

    @torch.no_grad()
    def input_video_stream(self, video_frames):
        self.session.push_video(video_frames)

    def input_query_stream(self, conversation):
        self.session.push_queries(conversation)

    @torch.no_grad()
    def _encode_frame(self):
        """returns: informative_score, relevance_score"""
        s = self.session
        if not s.frames:
            return None, None
        _, scores, _ = s.frame_pass(1)
        s.saw_frame()
        return {"informative_score": scores[0][0], "relevance_score": scores[0][1]}

    @torch.no_grad()
    def _encode_query(self):
        self.session.query_turn(self.session.queries.popleft()[1])

    @torch.no_grad()
    def _generate_response(self):
        return self.session.respond()

    @torch.no_grad()
    def inference(self):
        s = self.session
        turns = [{'time': t, 'content': q, 'role': 'user'} for t, q in s.queries]

        def on_response(t, text):
            turns.append({'time': t, 'content': text, 'role': 'assistant'})
            self.consecutive_n_frames = 0
        while s.frames:
            with s.lock:
                s.advance(on_response)
        return sorted(turns, key=lambda x: x['time'])


class LiveInferForDemo(LiveInferForBenchmark):
    def encode_given_query(self, query):
        with self.session.lock:
            self.session.query_turn(query)

    @torch.no_grad()
    def input_one_frame(self):
        s = self.session
        with s.lock:
            _, scores, _ = s.frame_pass(1)
            s.saw_frame()
            sc = {"informative_score": scores[0][0], "relevance_score": scores[0][1]}
            ret = dict(frame_idx=s.n_frames_seen, time=round(s.clock, 1), **sc)
            ret['response'] = s.respond() if s.rule.observe(sc) else None
            if ret['response'] is not None:
                self.consecutive_n_frames = 0
            s.clock += 1 / s.fps
            return ret


def round_numbers(data, n):
    """test/inference.py:322-329: debug_data is rounded to n decimals when written."""
    if isinstance(data, dict):
        return {k: round_numbers(v, n) for k, v in data.items()}
    if isinstance(data, list):
        return [round_numbers(v, n) for v in data]
    return round(data, n) if isinstance(data, float) else data
