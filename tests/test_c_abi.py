"""The drop-in boundary is a C ABI: the header is valid C (CPU check) and a plain-C program linked against
libec_b200.so + libcudart reproduces the oracle's per-instance disagreement (GPU check)."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")
SRC = os.path.join(ROOT, "tests", "c_abi", "disagreement_smoke.c")


def test_header_is_plain_c(tmp_path):
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    probe = tmp_path / "probe.c"
    probe.write_text('#include "ec_b200.h"\nint main(void) { ec_config c; (void)c; return EC_OK; }\n')
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                    str(probe)], check=True)


def _build(tmp_path):
    from embodied_captioning_b200 import _native as nat
    exe = str(tmp_path / "disagreement_smoke")
    subprocess.run(["gcc", "-std=c99", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"), SRC,
                    "-o", exe, nat.LIB_PATH, "-L", os.path.join(CUDA, "lib64"), "-lcudart",
                    "-Wl,-rpath," + os.path.dirname(nat.LIB_PATH), "-Wl,-rpath," + os.path.join(CUDA, "lib64")], check=True)
    return exe


def test_c_program_links_against_the_library(tmp_path):
    """CPU: the C consumer compiles and links (every symbol it uses is exported with C linkage)."""
    if shutil.which("gcc") is None or not os.path.isfile(os.path.join(CUDA, "include", "cuda_runtime_api.h")):
        pytest.skip("no gcc / CUDA headers")
    from embodied_captioning_b200 import _native as nat
    if not os.path.isfile(nat.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    assert os.path.isfile(_build(tmp_path))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1])
def test_c_program_matches_oracle(tmp_path, mode):
    sys.path.insert(0, ROOT)
    from oracle import disagreement_oracle as do
    exe = _build(tmp_path)
    rng = np.random.default_rng(3)
    n, d, m = 300, 64, 17
    inst = rng.integers(0, m, size=n).astype(np.int64) * 1_000_003
    emb = (rng.normal(size=(m, d))[inst // 1_000_003] + 0.7 * rng.normal(size=(n, d))).astype(np.float32)
    text = f"{n} {d} {mode}\n" + " ".join(repr(float(v)) for v in emb.ravel()) + "\n" + " ".join(str(int(k)) for k in inst) + "\n"
    out = subprocess.run([exe], input=text, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    rows = [l.split() for l in out.stdout.strip().splitlines()]
    rk, rs, rc = do.instance_disagreement(emb, inst, mode="full" if mode == 0 else "pairs")
    assert [int(r[0]) for r in rows] == rk.tolist()
    assert [int(r[1]) for r in rows] == rc.tolist()
    got = np.array([float(r[2]) for r in rows])
    ok = ~np.isnan(rs)
    assert np.array_equal(np.isnan(got), np.isnan(rs)) and np.abs(got[ok] - rs[ok]).max() < 1e-3
