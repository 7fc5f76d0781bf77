"""CPU: the peer exchange's protocol (lpm_b200/csrc/lpmx_peer_protocol.h -- the code peer_push_kernel runs) compiled for the
host, every CUDA thread a std::thread (tests/cpp/peer_protocol_model.cpp): 2..8 ranks, ping-pong and same-buffer sequences
with slow readers, ragged / empty / odd segments, a rank that never shows up.  The second build runs the same model under
ThreadSanitizer: with the ready and done handshakes in place there is no data race on the exchanged buffers, i.e. the
happens-before chain (stores -> CTA barrier -> ticket -> last CTA -> release flag -> acquire -> reads) is complete.  Removing
either wait makes the model fail (checked by hand when the protocol was written).  The GPU's memory
model itself is outside what a host model can show."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "peer_protocol_model.cpp")
INC = os.path.join(ROOT, "lpm_b200", "csrc")


def _run(exe):
    p = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    return p.returncode, p.stdout, p.stderr


def test_peer_protocol_host_model(tmp_path):
    exe = str(tmp_path / "peer_model")
    subprocess.run(["g++", "-O1", "-std=c++20", "-pthread", "-I", INC, SRC, "-o", exe], check=True)
    rc, out, err = _run(exe)
    assert rc == 0, out + err
    assert out.split() == ["peer_of", "0", "exchange", "0", "missing", "0"]


def test_peer_protocol_host_model_has_no_data_race(tmp_path):
    exe = str(tmp_path / "peer_model_tsan")
    p = subprocess.run(["g++", "-O1", "-g", "-std=c++20", "-pthread", "-fsanitize=thread", "-Wno-tsan", "-I", INC, SRC, "-o", exe],
                       capture_output=True, text=True)
    if p.returncode != 0:
        pytest.skip("ThreadSanitizer runtime not available: " + p.stderr[-300:])
    rc, out, err = _run(exe)
    if "FATAL: ThreadSanitizer" in err and "data race" not in err:
        pytest.skip("ThreadSanitizer cannot run here: " + err[-300:])
    assert "data race" not in err, err[-3000:]
    assert rc == 0, out + err
