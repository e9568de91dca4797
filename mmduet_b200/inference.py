"""Frame loop of the reference, same class / method / attribute names and the same decision rule, over the CUDA model:
  LiveInferForBenchmark    test/inference.py:20-313   (reset, set_fps, input_video_stream, input_query_stream,
                                                       _encode_frame, _encode_query, _generate_response, inference)
  LiveInferForDemo         demo/liveinfer.py:60-105   (encode_given_query, input_one_frame)
Differences that do not change results: frame tokens stay on the device (the reference parks them on the CPU and copies
them back, test/inference.py:212,237); frame steps skip lm_head (its output is never read there); both scores come
back in ONE 8-byte device->host copy per frame instead of two .item() syncs."""
import collections
import math
import threading
from dataclasses import asdict

import torch

from .modeling_live import fast_greedy_generate


class LiveInferForBenchmark:
    def __init__(self, args, model=None, tokenizer=None) -> None:
        assert not (args.bf16 and args.fp16), "only one of --bf16 true and --fp16 true can be set"
        if not args.bf16:
            raise ValueError("the accelerated path computes in bf16 (run with --bf16 true, as every reference script does)")
        self.torch_dtype = torch.bfloat16
        if model is None:
            from . import build_model_and_tokenizer
            model, tokenizer = build_model_and_tokenizer(is_training=False, set_vision_inside=True, torch_dtype=self.torch_dtype, **asdict(args))
        self.model, self.tokenizer = model, tokenizer
        self.model.eval()
        self.image_processor = self.model.get_vision_tower().image_processor
        self.device = self.model.device

        # visual
        self.hidden_size = self.model.config.hidden_size
        if args.frame_fps > 0:
            self.set_fps(args.frame_fps)
        self.frame_resolution = self.model.config.frame_resolution
        self.frame_num_tokens = self.model.vision.tokens_per_frame
        self.frame_v_placeholder = self.model.config.v_placeholder * self.frame_num_tokens

        # generation
        self.system_prompt = args.system_prompt
        self.inplace_output_ids = torch.zeros(1, 200, device=self.device, dtype=torch.long)
        self.stream_end_prob_threshold = args.stream_end_prob_threshold
        self.response_min_interval_frames = args.response_min_interval_frames
        self.threshold_z = args.threshold_z
        self.first_n_frames_no_generate = args.first_n_frames_no_generate
        self.running_list_length = args.running_list_length
        self.stream_end_score_sum_threshold = args.stream_end_score_sum_threshold
        self.score_heads = args.score_heads.split(',')
        self.consecutive_n_frames_threshold = args.consecutive_n_frames_threshold

        if int(self.threshold_z is not None) + int(self.stream_end_prob_threshold is not None) + int(self.stream_end_score_sum_threshold is not None) != 1:
            raise ValueError(f'only one of --stream_end_prob_threshold, --threshold_z and --stream_end_score_sum_threshold can be set. However, they are: {self.stream_end_prob_threshold}, {self.threshold_z}, {self.stream_end_score_sum_threshold}')
        if self.threshold_z is not None and self.first_n_frames_no_generate is None:
            raise ValueError('--first_n_frames_no_generate must be set when --threshold_z is set')
        if self.threshold_z is not None:
            raise NotImplementedError('--threshold_z is only implemented by the reference\'s DEPRECATED _call_for_streaming loop')

        self.remove_assistant_turns = args.remove_assistant_turns
        self.eos_token_id = self.model.config.eos_token_id
        if self.eos_token_id is None:
            self.eos_token_id = getattr(self.tokenizer, "eos_token_id", None)
        self._start_ids = self.tokenizer.apply_chat_template([{'role': 'system', 'content': self.system_prompt}], return_tensors='pt').to(self.device)
        self._added_stream_prompt_ids = self.tokenizer.apply_chat_template([{}], add_stream_prompt=True, return_tensors='pt').to(self.device)
        self._added_stream_generation_ids = self.tokenizer.apply_chat_template([{}], add_stream_generation_prompt=True, return_tensors='pt').to(self.device)
        self.repetition_penalty = args.repetition_penalty
        self._lock = threading.Lock()   # the demo re-enters the step from another thread (demo/app.py:84-85)
        self.past_key_values = None
        self.reset()

    def set_fps(self, fps=None, frame_interval=None):
        assert fps is not None or frame_interval is not None
        assert not (fps is not None and frame_interval is not None)
        if fps is not None:
            self.frame_fps = fps
            self.frame_interval = 1 / self.frame_fps
        else:
            self.frame_interval = frame_interval
            self.frame_fps = 1 / self.frame_interval

    def reset(self):
        self.query_queue = collections.deque()
        self.frame_embeds_queue = collections.deque()
        self.video_time = 0
        self.frame_idx = 0
        self.last_role = 'system'
        self.video_tensor = None
        self.last_ids = torch.tensor([[]], device=self.device, dtype=torch.long)
        if self.past_key_values:
            self.past_key_values.storage.release()   # pages go back to the pool
        self.past_key_values = None
        self.debug_data_list = list()
        self.generated_token_ids = list()
        self.num_frames_no_reply = 0
        self.stream_end_prob_list = list()
        self.stream_end_score_sum = 0
        self.consecutive_n_frames = 0

    @torch.no_grad()
    def input_video_stream(self, video_frames):
        """video_frames: uint8 [T,3,384,384] (what test/datasets.py yields) or already-processed float pixel_values.
        uint8 frames are rescaled/normalised inside the patch-embed kernel instead of by the image processor."""
        video_frames = video_frames.to(self.device, non_blocking=True)
        batch_size = 32
        for batch_i in range(0, math.ceil(len(video_frames) / batch_size)):
            video_frames_batch = video_frames[batch_i * batch_size: batch_i * batch_size + batch_size]
            frame_embeds = self.model.visual_embed(video_frames_batch).split(self.frame_num_tokens)
            self.frame_embeds_queue.extend([((r + batch_i * batch_size) / self.frame_fps, f) for r, f in enumerate(frame_embeds)])

    def input_query_stream(self, conversation):
        for turn in conversation:
            if turn['role'] == 'user':
                self.query_queue.append((turn['time'], turn['content']))

    def _forward(self, *, ids, frames=None, lm="none", score="last"):
        view = self.past_key_values
        if not view:
            view = self.model.new_cache()
        out = self.model.decoder.step([dict(storage=view.storage, past=view.length, ids=ids, frames=frames)], score=score, lm=lm)
        self.past_key_values = out["views"][0]
        return out

    def _encode_frame(self):
        """returns: informative_score, relevance_score"""
        if not self.frame_embeds_queue:
            return None, None
        video_time, frame_embeds = self.frame_embeds_queue.popleft()
        if not self.past_key_values:
            self.last_ids = self._start_ids
        elif self.last_role == 'assistant' and not self.remove_assistant_turns:
            self.last_ids = torch.cat([self.last_ids, self._added_stream_prompt_ids], dim=1)
        else:       # last_role is stream, now we just input another frame
            self.last_ids = torch.tensor([[]], device=self.device, dtype=torch.long)
        out = self._forward(ids=self.last_ids.view(-1).tolist(), frames=frame_embeds.view(-1, self.hidden_size))
        self.frame_idx += 1
        self.num_frames_no_reply += 1
        informative_score, relevance_score = out["scores"][0].tolist()   # one D2H read for both heads
        self.last_role = 'stream'
        return {"informative_score": informative_score, "relevance_score": relevance_score}

    def _encode_query(self):
        query_time, query = self.query_queue.popleft()
        self.last_ids = self.tokenizer.apply_chat_template([{'role': 'user', 'content': query}], add_stream_query_prompt=self.last_role == 'stream', add_stream_prompt=True, return_tensors='pt').to(self.device)
        out = self._forward(ids=self.last_ids.view(-1).tolist(), lm="last", score="none")
        self.last_ids = out["lm_logits"][:, :].argmax(dim=-1).view(1, 1)
        self.last_role = 'user'

    def _generate_response(self):
        self.last_ids = self._added_stream_generation_ids
        inputs_embeds = self.model.get_input_embeddings()(self.last_ids)
        view0 = self.past_key_values
        output_ids, past_key_values, self.generated_token_ids = fast_greedy_generate(
            model=self.model, inputs_embeds=inputs_embeds, past_key_values=view0, eos_token_id=self.eos_token_id,
            inplace_output_ids=self.inplace_output_ids, repetition_penalty=self.repetition_penalty,
            generated_token_ids=self.generated_token_ids)
        if not self.remove_assistant_turns:
            self.past_key_values = past_key_values
            self.last_ids = output_ids[:, -1:]
        else:
            # the returned cache is dropped: the next step appends at view0.length, i.e. the context rolls back
            # (transformers 4.44.2 legacy-cache meaning of test/inference.py:265-269)
            self.last_ids = torch.tensor([[]], device=self.device, dtype=torch.long)
        response = self.tokenizer.decode(output_ids[0], skip_special_tokens=True, clean_up_tokenization_spaces=True)
        self.num_frames_no_reply = 0
        self.last_role = 'assistant'
        return response

    def _decide(self, video_scores):
        need_response = False
        stream_end_score = sum([v for k, v in video_scores.items() if k in self.score_heads])
        self.stream_end_prob_list.append(stream_end_score)
        self.stream_end_score_sum += stream_end_score
        if isinstance(self.running_list_length, int) and self.running_list_length > 0:
            self.stream_end_prob_list = self.stream_end_prob_list[-self.running_list_length:]
        if self.stream_end_score_sum_threshold is not None and self.stream_end_score_sum > self.stream_end_score_sum_threshold:
            need_response = True
            self.stream_end_score_sum = 0
        if self.stream_end_prob_threshold is not None and stream_end_score > self.stream_end_prob_threshold:
            need_response = True
        return need_response

    # ---- multi-frame decoder passes (frames_per_step > 1) --------------------------------------------------------
    # Causal attention makes a k-frame pass arithmetically identical to k single-frame steps (tests/test_oracle.py::
    # test_chunked_frames_equal_stepwise), so the weights are streamed once per k frames.  The sequential decision rule is
    # preserved: the heads are evaluated at every frame's last token, the first crossing frame j is found, the KV cache is
    # rolled back to the end of frame j (O(1)) and the unconsumed frames return to the queue.
    frames_per_step = 1

    def _chunk_len(self):
        k = min(self.frames_per_step, len(self.frame_embeds_queue))
        if self.query_queue:
            # a query is encoded before the first frame whose video_time >= query time (inference() step 1)
            t_query = self.query_queue[0][0]
            n = 0
            while n < k and not (n > 0 and self.video_time + n / self.frame_fps >= t_query):
                n += 1
            k = max(n, 1)
        return k

    def _encode_frames_chunk(self, k):
        frames = [self.frame_embeds_queue.popleft() for _ in range(k)]
        if not self.past_key_values:
            self.last_ids = self._start_ids
        elif self.last_role == 'assistant' and not self.remove_assistant_turns:
            self.last_ids = torch.cat([self.last_ids, self._added_stream_prompt_ids], dim=1)
        else:
            self.last_ids = torch.tensor([[]], device=self.device, dtype=torch.long)
        prefix = self.last_ids.view(-1).tolist()
        P, n = len(prefix), self.frame_num_tokens
        past = self.past_key_values.length if self.past_key_values else 0
        view = self.past_key_values if self.past_key_values else self.model.new_cache()
        emb = torch.cat([f[1].view(-1, self.hidden_size) for f in frames], 0)
        out = self.model.decoder.step([dict(storage=view.storage, past=view.length, ids=prefix, frames=emb,
                                            score_rows=[P + n * (j + 1) - 1 for j in range(k)])], score="frame_ends")
        self.past_key_values = out["views"][0]
        scores = out["scores"].tolist()                               # one D2H read for the whole chunk
        lens = [past + P + n * (j + 1) for j in range(k)]
        return frames, scores, lens

    @torch.no_grad()
    def _inference_chunked(self, model_response_list):
        from .engine import CacheView
        while self.frame_embeds_queue:
            with self._lock:
                if self.query_queue and self.video_time >= self.query_queue[0][0]:
                    self._encode_query()
                k = self._chunk_len()
                frames, scores, lens = self._encode_frames_chunk(k)
                for j in range(k):
                    self.frame_idx += 1
                    self.num_frames_no_reply += 1
                    self.last_role = 'stream'
                    video_scores = {"informative_score": scores[j][0], "relevance_score": scores[j][1]}
                    self.debug_data_list.append(dict(time=self.video_time, **video_scores))
                    if self._decide(video_scores):
                        if j + 1 < k:   # speculative frames j+1.. are undone: KV rollback + back to the queue
                            self.past_key_values = CacheView(self.past_key_values.storage, lens[j])
                            self.past_key_values.storage.truncate(lens[j])
                            for f in reversed(frames[j + 1:]):
                                self.frame_embeds_queue.appendleft(f)
                        response = self._generate_response()
                        model_response_list.append({'time': self.video_time, 'content': response, 'role': 'assistant'})
                        self.num_frames_no_reply = 0
                        self.consecutive_n_frames = 0
                        self.video_time += 1 / self.frame_fps
                        break
                    self.video_time += 1 / self.frame_fps
        return sorted(model_response_list, key=lambda x: x['time'])

    @torch.no_grad()
    def inference(self):
        model_response_list = [{'time': q[0], 'content': q[1], 'role': 'user'} for q in self.query_queue]
        if self.frames_per_step > 1:
            return self._inference_chunked(model_response_list)
        while self.frame_embeds_queue:
            with self._lock:
                # 1. check if a user query is at current time
                if self.query_queue and self.video_time >= self.query_queue[0][0]:
                    self._encode_query()
                # 2. input a frame, and update the scores list
                video_scores = self._encode_frame()
                self.debug_data_list.append(dict(time=self.video_time, **video_scores))
                # 3. check the scores, if need to generate a response
                need_response = self._decide(video_scores)
                # 4. record the responses
                if need_response:
                    response = self._generate_response()
                    model_response_list.append({'time': self.video_time, 'content': response, 'role': 'assistant'})
                    self.num_frames_no_reply = 0
                    self.consecutive_n_frames = 0
                # 5. update the video time
                self.video_time += 1 / self.frame_fps
        return sorted(model_response_list, key=lambda x: x['time'])


class LiveInferForDemo(LiveInferForBenchmark):
    def encode_given_query(self, query):
        with self._lock:
            self.last_ids = self.tokenizer.apply_chat_template([{'role': 'user', 'content': query}], add_stream_query_prompt=self.last_role == 'stream', add_stream_prompt=True, return_tensors='pt').to(self.device)
            out = self._forward(ids=self.last_ids.view(-1).tolist(), lm="last", score="none")
            self.last_ids = out["lm_logits"].argmax(dim=-1).view(1, 1)
            self.last_role = 'user'

    @torch.no_grad()
    def input_one_frame(self):
        with self._lock:
            video_scores = self._encode_frame()
            ret = dict(frame_idx=self.frame_idx, time=round(self.video_time, 1), **video_scores)
            need_response = self._decide(video_scores)
            if need_response:
                response = self._generate_response()
                self.num_frames_no_reply = 0
                self.consecutive_n_frames = 0
            else:
                response = None
            ret['response'] = response
            self.video_time += 1 / self.frame_fps
            return ret


def round_numbers(data, n):
    """test/inference.py:322-329 (debug_data is rounded to 3 decimals when written)."""
    if isinstance(data, list):
        return [round_numbers(d, n) for d in data]
    elif isinstance(data, dict):
        return {k: round_numbers(v, n) for k, v in data.items()}
    elif isinstance(data, float):
        return round(data, n)
    return data
