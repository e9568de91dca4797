"""The driver loop of the reference's `python -m test.inference` (test/inference.py:332-361) over the CUDA path: for every
item (question_id, uint8 frames, conversation, fps, duration) run the frame loop and write one JSON line per video with
the reference's schema: question_id, model_response_list, video_duration, debug_data (rounded to 3 decimals), so that
test/evaluate.py keeps working on the output.  Video decoding (cv2, test/datasets.py) stays outside:
any iterable of already-decoded items can be passed."""
import json

from .inference import LiveInferForBenchmark, round_numbers


def run(infer: LiveInferForBenchmark, items, output_fname, grounding_mode=False, log=None):
    if grounding_mode:
        infer.first_n_frames_no_generate = 100000   # kept for parity with test/inference.py:365 (unused by the live loop)
    n = 0
    with open(output_fname, "w") as f_out:
        for data_i, data in enumerate(items):
            question_id, video_frames, conversation, fps, video_duration = data
            if question_id is None:
                continue
            infer.reset()
            if log:
                log(f"num frames and fps for {question_id}: {len(video_frames)}, {fps}")
            infer.set_fps(fps=fps)
            infer.input_video_stream(video_frames)
            infer.input_query_stream(conversation)
            model_response_list = infer.inference()
            res = {"question_id": question_id, "model_response_list": model_response_list, "video_duration": video_duration}
            res["debug_data"] = round_numbers(infer.debug_data_list, 3)
            f_out.write(json.dumps(res) + "\n")
            if data_i % 5 == 0:
                f_out.flush()
            n += 1
    return n
