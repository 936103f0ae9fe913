#!/usr/bin/env python
"""Write the canonical float32 weight blob `cvad_create` takes (order: DESIGN.md section 2), read from the reference's own
.onnx file -- for hosts that use the C ABI without Python (examples/c_abi_demo.c).

    python tools/export_weights.py v5|v4|v4_8k out.f32 [model.onnx]
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "cutter-vad_b200"))
from real_time_vad.engine.onnx_weights import canonical_blob_v4, canonical_blob_v5  # noqa: E402
from real_time_vad.engine.stream_engine import MODELS_DIR  # noqa: E402


def main():
    kind, out = sys.argv[1], sys.argv[2]
    if kind == "v5":
        blob = canonical_blob_v5(Path(sys.argv[3]) if len(sys.argv) > 3 else MODELS_DIR / "silero_vad_v5.onnx")
    else:
        path = Path(sys.argv[3]) if len(sys.argv) > 3 else MODELS_DIR / "silero_vad.onnx"
        blob = canonical_blob_v4(path, branch="8k") if kind == "v4_8k" else canonical_blob_v4(path)
    blob.tofile(out)
    print(f"{out}: {blob.size} float32")


if __name__ == "__main__":
    main()
