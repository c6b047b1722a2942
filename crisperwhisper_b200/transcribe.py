"""CLI with the reference's interface (REF/transcribe.py): `python -m crisperwhisper_b200.transcribe --f audio.wav`.

Loads the checkpoint named by --model_id with HF's loaders (weights only — REF/transcribe.py:12-19), repacks it for
libcrisper.so and runs the B200 pipeline.  WAV input (the reference shells out to ffmpeg, which is outside this path)."""
import argparse
import os
import sys


def transcribe_audio(file_path, model_id="nyrahealth/CrisperWhisper"):
    import torch
    from transformers import AutoModelForSpeechSeq2Seq, AutoProcessor
    from .asr_pipeline import pipeline
    from .utils import adjust_pauses_for_hf_pipeline_output

    if not torch.cuda.is_available():
        raise RuntimeError("crisperwhisper_b200 needs a B200 (sm_100a) GPU; there is no CPU path")
    model = AutoModelForSpeechSeq2Seq.from_pretrained(model_id, torch_dtype=torch.bfloat16, low_cpu_mem_usage=False,
                                                      use_safetensors=True)
    processor = AutoProcessor.from_pretrained(model_id)
    pipe = pipeline("automatic-speech-recognition", model=model, tokenizer=processor.tokenizer,
                    feature_extractor=processor.feature_extractor, chunk_length_s=30, batch_size=16,
                    return_timestamps="word", device="cuda:0")
    return adjust_pauses_for_hf_pipeline_output(pipe(file_path))


def main():
    parser = argparse.ArgumentParser(description="Transcribe an audio file.")
    parser.add_argument("--f", type=str, required=True, help="Path to the audio file")
    parser.add_argument("--model_id", type=str, default="nyrahealth/CrisperWhisper")
    parser.add_argument("--vtt", type=str, default=None, help="also write the word timestamps as WebVTT (REF/app.py:74-82)")
    args = parser.parse_args()
    if not os.path.exists(args.f):
        print(f"Error: The file '{args.f}' does not exist.")
        sys.exit(1)
    try:
        transcription = transcribe_audio(args.f, args.model_id)
        print("Transcription:")
        print(transcription["text"])
        if args.vtt:
            from .utils import timestamps_to_vtt
            with open(args.vtt, "w", encoding="utf-8") as f:
                f.write(timestamps_to_vtt(transcription["chunks"]))
    except Exception as e:  # same error contract as the reference CLI (REF/transcribe.py:46-52)
        print(f"An error occurred while transcribing the audio: {str(e)}")
        sys.exit(1)


if __name__ == "__main__":
    main()
