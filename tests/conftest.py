import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the oracle (test infra) if missing; the CUDA library is built by __graft_entry__.build()."""
    so = os.path.join(ROOT, "oracle", "_build", "liborc.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    yield


@pytest.fixture(scope="session", autouse=True)
def _warm_kernel_cache(request, _built):
    """On a GPU box, before the first GPU test: compile the plan-specialised kernels the parity tests use that are not
    in the on-disk cache yet, on all host cores at once (tests/jit_warm.py).  One after the other inside the tests they
    cost 5-100 s of NVRTC time each; results do not depend on this step."""
    def have_gpu():
        if os.path.exists("/dev/nvidia0") or os.path.exists("/dev/nvidiactl"):
            return True
        try:
            return "GPU" in subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=20).stdout
        except Exception:
            return False
    if "not gpu" not in (request.config.getoption("-m") or "") and not os.environ.get("B2_NO_JIT_WARM") and have_gpu():
        import time
        import jit_warm
        t0 = time.time()
        done, errs = jit_warm.warm()
        try:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "jit_warm.log"), "w") as f:
                f.write(f"warm-up: {done} kernels compiled in {time.time() - t0:.0f} s, errors: {errs[:2]}\n")
        except Exception:
            pass
        yield
        try:  # how many kernels the tests still had to compile themselves (tuning aid, never a failure)
            import ctypes as C
            from tikv_b200 import ffi
            a, b = C.c_uint64(), C.c_uint64()
            ffi.lib().b2_jit_counters(C.byref(a), C.byref(b))
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "jit_warm.log"), "a") as f:
                f.write(f"during the tests: {a.value} NVRTC compilations, {b.value} disk-cache hits\n")
        except Exception:
            pass
        return
    yield
