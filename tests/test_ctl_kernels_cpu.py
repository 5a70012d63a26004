"""The control kernels of the byte stages and of the receiver's seam resolution, run on the HOST.

`leansdr_b200/csrc/k_ctl_fec.cuh` / `k_ctl_rx.cuh` hold device code only; nvcc compiles them into the library, and
here g++ compiles the same text against `tests/emu/cuda_emu.h` (one OS thread per CUDA thread, barriers for
`__syncthreads`, exchange slots for shuffles and ballots).  `tests/emu/emu_ctl.cpp` feeds every kernel and its
predecessor (`tests/emu/ctl_v1.cuh`: the kernels as they passed the GPU parity suite on B200) the same seeded inputs
and compares every output word.  This is how a rewrite of these integer programs is checked where there is no GPU;
the GPU parity tests then check them again against the oracle through the C ABI.
"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"


@pytest.fixture(scope="module")
def emu_bin(tmp_path_factory):
    if shutil.which("g++") is None or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("g++ or the CUDA headers are not available")
    out = str(tmp_path_factory.mktemp("emu") / "emu_ctl")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-pthread", "-w", "-I", CUDA_INC,
                           os.path.join(ROOT, "tests", "emu", "emu_ctl.cpp"), "-o", out])
    return out


@pytest.mark.timeout(900)
@pytest.mark.parametrize("case,seeds", [("plan", (1,)), ("derand", (1,)), ("sync_locked", (1, 2)),
                                        ("sync_search", (1,)), ("deconv", (1,))])
def test_control_kernel_equals_its_predecessor(emu_bin, case, seeds):
    for seed in seeds:
        r = subprocess.run([emu_bin, case, str(seed)], capture_output=True, text=True, timeout=800)
        assert r.returncode == 0, f"{case} seed {seed}:\n{r.stderr[-2000:]}"
        assert "identical" in r.stdout


ORACLE_LIB = os.path.join(ROOT, "oracle", "liboracle.so")
EMU_VIT_SRC = [os.path.join(ROOT, "tests", "emu", "emu_vit.cpp"), os.path.join(ROOT, "leansdr_b200", "csrc", "tables.cpp"),
               ORACLE_LIB, "-Wl,-rpath," + os.path.dirname(ORACLE_LIB)]


@pytest.fixture(scope="module")
def emu_vit_bin(tmp_path_factory, oracle):
    if shutil.which("g++") is None or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("g++ or the CUDA headers are not available")
    out = str(tmp_path_factory.mktemp("emu") / "emu_vit")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-pthread", "-w", "-I", CUDA_INC, *EMU_VIT_SRC, "-o", out])
    return out


# (fec, kernel, seed, stream, layout): fec 5 = 7/8 through k_viterbi<kVitFull> (the rescan replaced by the minimum of the
# metrics + the tie rule), fec 0 = 1/2 through k_viterbi_ws (one warp per time segment); streams with a decodable
# hypothesis and pure noise; segments far from / close to the start of the batch, and resync_period 1 (--fastlock).
@pytest.mark.timeout(900)
@pytest.mark.parametrize("args", [(5, "full", 1, "signal", 1), (5, "full", 3, "noise", 0), (0, "ws", 2, "noise", 1)])
def test_viterbi_kernel_equals_its_predecessor(emu_vit_bin, args):
    """k_vit_dev.cuh (the text nvcc compiles) on the host: output bytes, entry / exit states of every time segment and
    the elected hypothesis equal those of the kernel that passed the GPU parity suite (tests/emu/vit_v1.cuh), on cold
    segments, exact segments and the repair path."""
    r = subprocess.run([emu_vit_bin, *map(str, args)], capture_output=True, text=True, timeout=800)
    assert r.returncode == 0, f"{args}:\n{r.stderr[-2000:]}"
    assert "identical" in r.stdout


# (fec, kernel, seed, stream, layout): 0 = 1/2, 2 = 4/6, 3 = 3/4, 4 = 5/6, 5 = 7/8; layout 2 = resync_period 1 (--fastlock)
@pytest.mark.timeout(900)
@pytest.mark.parametrize("args", [(0, "ws", 1, "signal", 0), (0, "ws", 2, "noise", 2), (0, "r12", 1, "signal", 0),
                                  (5, "full", 1, "signal", 0), (5, "full", 2, "noise", 2), (5, "generic", 1, "signal", 0),
                                  (2, "generic", 1, "signal", 0), (3, "generic", 1, "signal", 0), (4, "generic", 2, "noise", 0)])
def test_viterbi_kernel_equals_the_oracle_on_the_host(emu_vit_bin, args):
    """The Viterbi kernels (the text nvcc compiles) against the ORACLE's viterbi_sync, which is pinned to the
    reference's: every hypothesis of the configuration (4 at QPSK 1/2, 16 at 7/8), one serial segment from the
    constructor state; output bytes, the 64 metrics and 64 path registers of every decoder at the end, the elected
    hypothesis and the re-sync phase are the oracle's -- on a decodable stream and on noise, for all five trellises
    QPSK takes, with resync_period 32-style voting and with a vote per chunk."""
    r = subprocess.run([emu_vit_bin, *map(str, args), "oracle"], capture_output=True, text=True, timeout=800)
    assert r.returncode == 0, f"{args}:\n{r.stderr[-2000:]}"
    assert "identical to the oracle" in r.stdout


def _tsan_build(tmp_path_factory, name, sources):
    if shutil.which("g++") is None or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("g++ or the CUDA headers are not available")
    out = str(tmp_path_factory.mktemp("tsan") / name)
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-g", "-fsanitize=thread", "-pthread", "-w", "-I", CUDA_INC, *sources, "-o", out],
                       capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("g++ cannot link ThreadSanitizer here: " + r.stderr[-300:])
    return out


def _tsan_warnings(cmd, timeout):
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout,
                       env=dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=0"))
    return r, (r.stdout + r.stderr).count("WARNING: ThreadSanitizer")


@pytest.mark.timeout(3600)
def test_emulated_kernels_are_race_free_under_thread_sanitizer(tmp_path_factory, oracle):
    """The host shim runs one OS thread per CUDA thread with real barriers behind __syncthreads / __syncwarp / the warp
    collectives, so ThreadSanitizer sees a shared-memory exchange that lacks one of them as a data race -- the class of
    bug a GPU hides until the warp scheduler changes.  First the detector is shown to work (a kernel with a missing
    __syncthreads and one with a missing __syncwarp are reported, their fixed versions are clean), then the Viterbi
    kernels of rate 1/2 (a warp per segment) and 7/8 (no rescan walk) run against the oracle without a report.  LDVB_EMU_TSAN=1 runs every
    emulated kernel this way (Viterbi generic / rate 1/2 / full trellis / warp per segment, seam plan, de-randomiser,
    lock tracker, deconvolution tiles: a few minutes; all clean at the end of round 2)."""
    emu = os.path.join(ROOT, "tests", "emu")
    sanity = _tsan_build(tmp_path_factory, "tsan_sanity", [os.path.join(emu, "tsan_sanity.cpp")])
    r, n = _tsan_warnings([sanity], 120)
    assert r.returncode == 0 and n >= 2, f"ThreadSanitizer did not report the seeded races ({n}):\n{r.stderr[-1500:]}"
    r, n = _tsan_warnings([sanity, "sync"], 120)
    assert r.returncode == 0 and n == 0, r.stderr[-1500:]
    vit = _tsan_build(tmp_path_factory, "emu_vit_tsan", EMU_VIT_SRC)
    cases = [["0", "ws", "1", "signal", "0", "oracle"], ["5", "full", "1", "signal", "0", "oracle"]]
    ctl_cases = []
    if os.environ.get("LDVB_EMU_TSAN") == "1":
        cases += [["0", "ws", "2", "noise", "1"], ["5", "full", "1", "signal", "1"], ["5", "full", "3", "noise", "2"],
                  ["3", "generic", "1", "signal", "0"], ["0", "r12", "1", "signal", "0"]]
        ctl_cases = ["plan", "derand", "sync_locked", "sync_search", "deconv"]
    for c in cases:
        r, n = _tsan_warnings([vit, *c], 800)
        assert r.returncode == 0 and "identical" in r.stdout and n == 0, f"{c}: {n} reports\n{(r.stdout + r.stderr)[-3000:]}"
    if ctl_cases:
        ctl = _tsan_build(tmp_path_factory, "emu_ctl_tsan", [os.path.join(emu, "emu_ctl.cpp")])
        for c in ctl_cases:
            r, n = _tsan_warnings([ctl, c, "1"], 2400)
            assert r.returncode == 0 and "identical" in r.stdout and n == 0, f"{c}: {n} reports\n{(r.stdout + r.stderr)[-3000:]}"
