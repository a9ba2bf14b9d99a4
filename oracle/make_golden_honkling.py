"""TEST INFRASTRUCTURE ONLY -- add the Honkling-export known answer to ``tests/golden/meta.json``.

    python oracle/make_golden_honkling.py

Runs the reference's own ``training/run/export_honkling.py`` (unmodified, via ``runpy``) on the shipped hey-fire-fox res8
checkpoint and records the SHA-256, size and first bytes of the file it writes.  ``howl_b200.export.export_honkling``
must reproduce that file byte for byte from the same state dict (``tests/golden/res8_heyfirefox.npz`` holds it).
"""
import hashlib
import json
import os
import runpy
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import OUT, REF, _install_shims  # noqa: E402


def main():
    _install_shims()
    ckpt = os.path.join(REF, "howl-models", "howl", "hey-fire-fox", "model-best.pt.bin")
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "honkling.js")
        argv = sys.argv
        sys.argv = ["export_honkling.py", "--input-file", ckpt, "--output-file", out, "--name", "RES8"]
        import torch

        real_load = torch.load
        # the shipped checkpoint holds CUDA tensors and the script calls torch.load(path) bare: map to CPU in this GPU-less container
        torch.load = lambda f, *a, **k: real_load(f, *a, **{"map_location": "cpu", **k})
        try:
            runpy.run_path(os.path.join(REF, "training", "run", "export_honkling.py"), run_name="__main__")
        finally:
            sys.argv = argv
            torch.load = real_load
        data = open(out, "rb").read()
    meta_path = os.path.join(OUT, "meta.json")
    meta = json.load(open(meta_path))
    meta["honkling"] = {"name": "RES8", "sha256": hashlib.sha256(data).hexdigest(), "bytes": len(data),
                        "head": data[:96].decode(), "tail": data[-48:].decode()}
    json.dump(meta, open(meta_path, "w"), indent=1, sort_keys=True)
    print(meta["honkling"])


if __name__ == "__main__":
    main()
