"""Flag names and defaults of the reference's inference CLI (models/arguments_live.py:5-55), as a plain dataclass
(the reference derives from transformers.TrainingArguments; only the fields the frame loop reads are kept, plus
bf16/fp16 which LiveInferForBenchmark.__init__ asserts on)."""
from dataclasses import dataclass, field


@dataclass
class LiveTestArguments:
    system_prompt: str = (
        "A multimodal AI assistant is helping users with some activities."
        " Below is their conversation, interleaved with the list of video frames received by the assistant."
    )
    live_version: str = "test"
    llm_pretrained: str = "lmms-lab/llava-onevision-qwen2-7b-ov"
    vision_pretrained: str = "google/siglip-large-patch16-384"
    lora_pretrained: str = None
    attn_implementation: str = "flash_attention_2"   # accepted and ignored: there is a single (CUDA) backend
    bf16: bool = True
    fp16: bool = False
    frame_fps: float = 2
    frame_token_cls: bool = False
    frame_token_pooled: list = field(default_factory=lambda: [7, 7])
    frame_num_tokens: int = 49
    video_pooling_stride: int = 4
    frame_resolution: int = 384
    v_placeholder: str = "<image>"
    max_num_frames: int = 100
    is_online_model: bool = True
    grounding_mode: bool = False
    input_dir: str = "datasets/shot2story/videos/"
    test_fname: str = ""
    output_fname: str = ""
    repetition_penalty: float = None
    stream_end_prob_threshold: float = None
    response_min_interval_frames: int = None
    threshold_z: float = None
    first_n_frames_no_generate: int = 0
    consecutive_n_frames_threshold: int = 1
    running_list_length: int = 20
    start_idx: int = 0
    end_idx: int = None
    time_instruction_format: str = None
    stream_end_score_sum_threshold: float = None
    remove_assistant_turns: bool = False
    score_heads: str = "informative_score"
