"""CLI with the reference's interface (REF/transcribe.py): `python -m crisperwhisper_b200.transcribe --f audio.wav`.

Loads the checkpoint named by --model_id with HF's loaders (weights only — REF/transcribe.py:12-19), repacks it for
libcrisper.so and runs the B200 pipeline.  WAV input (the reference shells out to ffmpeg, which is outside this path)."""
import argparse
import os
import sys


def transcribe_audio(file_path, model_id="nyrahealth/CrisperWhisper", adjust_pauses=False):
    import torch
    from transformers import AutoModelForSpeechSeq2Seq, AutoProcessor
    from .asr_pipeline import pipeline
    from .utils import adjust_pauses_for_hf_pipeline_output

    if not torch.cuda.is_available():
        raise RuntimeError("crisperwhisper_b200 needs a B200 (sm_100a) GPU; there is no CPU path")
    # float32 on the host: pack_state_dict does the single bf16 rounding of the matrices, and LayerNorm gamma/beta, biases
    # and the positional tables reach their f32 slots unrounded.
    model = AutoModelForSpeechSeq2Seq.from_pretrained(model_id, torch_dtype=torch.float32, low_cpu_mem_usage=False,
                                                      use_safetensors=True)
    processor = AutoProcessor.from_pretrained(model_id)
    pipe = pipeline("automatic-speech-recognition", model=model, tokenizer=processor.tokenizer,
                    feature_extractor=processor.feature_extractor, chunk_length_s=30, batch_size=16,
                    return_timestamps="word", device="cuda:0")
    out = pipe(file_path)
    # The reference CLI prints the raw pipeline output (REF/transcribe.py:33-35); the pause redistribution is the
    # README's optional post-step (REF/README.md:174, REF/utils.py:1) — opt in with --adjust_pauses.
    return adjust_pauses_for_hf_pipeline_output(out) if adjust_pauses else out


def main():
    parser = argparse.ArgumentParser(description="Transcribe an audio file.")
    parser.add_argument("--f", type=str, required=True, help="Path to the audio file")
    parser.add_argument("--model_id", type=str, default="nyrahealth/CrisperWhisper")
    parser.add_argument("--vtt", type=str, default=None, help="also write the word timestamps as WebVTT (REF/app.py:74-82)")
    parser.add_argument("--adjust_pauses", action="store_true",
                        help="redistribute pauses between words (REF/utils.py:1 adjust_pauses_for_hf_pipeline_output, REF/README.md:174)")
    args = parser.parse_args()
    if not os.path.exists(args.f):
        print(f"Error: The file '{args.f}' does not exist.")
        sys.exit(1)
    try:
        transcription = transcribe_audio(args.f, args.model_id, args.adjust_pauses)
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
