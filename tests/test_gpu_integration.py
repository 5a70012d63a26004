"""Drop-in check: the reference's OWN scheduler / pipebuf / file_reader / file_writer
(compiled from /root/reference in the dev container into oracle/_ref/leandvb_gpu, together
with leansdr_b200/host/gpu_runnables.h) driving the CUDA path, against the unmodified
reference `leandvb` on the same IQ."""
import os
import subprocess

import numpy as np
import pytest

from tests import vectors as V

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "leandvb_gpu")),
                    reason="oracle/_ref/leandvb_gpu not built (needs /root/reference at build time)")
@pytest.mark.parametrize("fmt,flags", [("f32", ["--resample"]), ("u8", []), ("f32", ["--anf", "0", "--gpu-exact"]),
                                       ("u8", ["--hs", "--gpu-exact"]), ("u8", ["--hs"]), ("f32", ["--fastlock"])])
def test_reference_scheduler_runs_gpu_runnable(product, oracle, fmt, flags):
    O = oracle
    raw = V.ref_iq(1500, fmt=fmt)
    base = ["--" + fmt, "-f", "2400e3", "--sr", "2000e3", "--cr", "1/2"]
    ref_flags = [f for f in flags if not f.startswith("--gpu")]
    want = V.ref_leandvb(raw, base + ref_flags)
    out = subprocess.run([O.ref_bin("leandvb_gpu"), *base, *flags, "--gpu-batch", str(1 << 20)],
                         input=raw.tobytes(), stdout=subprocess.PIPE, check=True, timeout=120).stdout
    got = np.frombuffer(out, dtype=np.uint8).reshape(-1, 188)
    n = min(len(got), len(want))
    assert n > 1400
    assert np.array_equal(got[:n], want[:n])
    assert 0 <= len(got) - len(want) <= 1


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "leandvbtx_gpu")),
                    reason="oracle/_ref/leandvbtx_gpu not built (needs /root/reference at build time)")
@pytest.mark.parametrize("flags,batch", [(["-f", "6/5", "--power", "37.5", "--agc"], 256),
                                         (["--cr", "7/8", "-f", "2", "--power", "37.5", "--agc"], 4096),
                                         (["--const", "8PSK", "--cr", "2/3", "-f", "4"], 100)])
def test_reference_scheduler_runs_gpu_transmitter(product, oracle, flags, batch):
    """leantsgen | leandvbtx_gpu (the reference's scheduler driving gpu_dvbs_transmitter) equals
    leantsgen | leandvbtx (the unmodified reference), bit for bit including the length."""
    O = oracle
    ts = subprocess.run([O.ref_bin("leantsgen"), "-c", "1500"], stdout=subprocess.PIPE, check=True).stdout
    want = subprocess.run([O.ref_bin("leandvbtx"), *flags], input=ts, stdout=subprocess.PIPE, check=True).stdout
    got = subprocess.run([O.ref_bin("leandvbtx_gpu"), *flags, "--gpu-batch", str(batch)], input=ts,
                         stdout=subprocess.PIPE, check=True, timeout=120).stdout
    assert len(got) == len(want) and len(got) > 1000000
    assert got == want
