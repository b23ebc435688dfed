"""The C++ host mirror (arrow-rs_b200/host/arrow_cuda.hpp): builds on CPU, and on a GPU its test
binary — the reference's unit tests re-expressed in C++ — must pass."""
import os
import subprocess

import pytest

from acu import _abi as abi

HOST = os.path.join(abi.REPO, "arrow-rs_b200", "host")
BIN = os.path.join(HOST, "test_host")


def test_host_mirror_builds():
    subprocess.run(["make", "-s", "-C", HOST], check=True)
    assert os.path.exists(BIN)


def test_host_binary_refuses_to_run_without_gpu():
    import ctypes as C
    lib = abi.load_library()
    h = C.c_void_p()
    if lib.acu_ctx_create(0, C.byref(h)) == abi.OK:
        lib.acu_ctx_destroy(h)
        pytest.skip("CUDA device present")
    r = subprocess.run([BIN], capture_output=True, text=True)
    assert r.returncode == 77 and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_host_mirror_reference_tests_pass():
    if not os.path.exists(BIN):
        subprocess.run(["make", "-s", "-C", HOST], check=True)
    env = dict(os.environ)
    try:  # an IPC stream written by another Arrow implementation for the StreamReader test
        import io
        import tempfile
        import pyarrow as pa
        t = pa.table({"a": pa.array([1, None, 3], type=pa.int64()), "s": pa.array(["x", None, "hello"]),
                      "b": pa.array([True, False, None]), "f": pa.array([1.5, 2.5, 3.5], type=pa.float32())})
        sink = io.BytesIO()
        with pa.ipc.new_stream(sink, t.schema) as w:
            w.write_table(t)
            w.write_table(t.slice(1, 2))
        f = tempfile.NamedTemporaryFile(suffix=".arrows", delete=False)
        f.write(sink.getvalue())
        f.close()
        env["ACU_TEST_IPC_FILE"] = f.name
    except ImportError:
        pass
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=300, env=env)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
