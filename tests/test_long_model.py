"""The segmented front end of the long-stream encoder (lzfse_rust_b200/csrc/encode_long.cuh) as a CPU model:
tests/model/long_parse_model.c computes the per-position find words, replays every segment speculatively from a
clean state, stitches the segments in order and compares the resulting match list with the oracle's sequential
front end (orc_frontend_lmds).  This pins the ALGORITHM; the CUDA kernels are compared frame by frame on the GPU
(tests/test_gpu_configs.py::test_large_streams, test_long_encoder_assorted)."""
import os
import subprocess

import numpy as np
import pytest

import testkit as tk

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def model(tmp_path_factory):
    import oracle_binding  # noqa: F401  (builds oracle/liblzfse_oracle.so)

    exe = str(tmp_path_factory.mktemp("model") / "long_parse_model")
    subprocess.run(["gcc", "-O2", "-o", exe, os.path.join(HERE, "model", "long_parse_model.c"), "-L" + os.path.join(ROOT, "oracle"), "-llzfse_oracle",
                    "-Wl,-rpath," + os.path.join(ROOT, "oracle")], check=True)
    return exe


def _inputs():
    rng = np.random.default_rng(1)
    t = tk.synth_text(9, 400000)
    r = bytearray(rng.integers(0, 256, 300000, dtype=np.uint8).tobytes())
    for _ in range(600):  # sparse short repeats in noise: matches stay pending across segment borders
        a = int(rng.integers(100000, len(r) - 100)); d = int(rng.integers(8, 90000)); n = int(rng.integers(4, 30))
        r[a:a + n] = r[a - d:a - d + n]
    mix = bytearray()
    while len(mix) < 300000:
        k = int(rng.integers(0, 4)); n = int(rng.integers(1, 2000))
        if k == 0: mix += t[int(rng.integers(0, len(t) - n * 10)):][:n * 10]
        elif k == 1: mix += bytes([int(rng.integers(0, 256))]) * min(n, 300)
        elif k == 2: mix += rng.integers(0, 256, n * 5, dtype=np.uint8).tobytes()
        else: mix += (rng.integers(0, 4, n * 5, dtype=np.uint8) * 16).tobytes()
    far = tk.rng_gen_vec(3, 90000)
    return {"far_repeat": far + far + t[:30000] + far[:50000], "text": t, "noise": tk.rng_gen_vec(3, 200000), "sparse": bytes(r), "mix": bytes(mix), "zeros": bytes(70000)}


@pytest.mark.parametrize("seg", [4096, 16384, 65536])
def test_segmented_parse_equals_sequential(model, tmp_path, seg):
    for name, data in _inputs().items():
        f = tmp_path / name
        f.write_bytes(data)
        r = subprocess.run([model, str(f), str(seg)], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0 and r.stdout.startswith("OK"), (name, seg, r.stdout)
