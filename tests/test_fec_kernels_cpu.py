"""The packet kernels of `leansdr_b200/csrc/k_fec.cu` run on the HOST and checked against the oracle.

The device text of the file (everything inside its anonymous namespace: the launchers stay behind) is cut out here and
compiled by g++ against `tests/emu/cuda_emu.h` together with `tests/emu/emu_fec.cpp`: the Reed-Solomon decoder with
and without the fused de-interleaver gather (0..10 byte errors per packet: flags, corrected-bit counts and bytes equal
the oracle's `rs_decoder`), byte re-alignment and sync flags (equal mpeg_sync's definition), the de-randomiser's
scan + output grids (equal the oracle's `derandomizer`, dropped packets and carried position included), and the
algebraic deconvolver (`k_deconv_tiled` + the carry thread: bytes, symbols consumed, carried shift register and leftover
bits equal the oracle's `deconvol_sync` for every code rate, each of the four hypotheses, over two batches), and the
MPEG sync tracker (`k_sync_flags` + `k_sync_track` + `k_realign`, driven pass by pass like `run_sync` of pipeline.cu:
bytes consumed and produced, lock state, `next_sync` requests and aligned bytes equal the oracle's `mpeg_sync` on
streams with a bit offset, either polarity, garbage in front, a burst that loses the lock and a re-acquisition), and the
hard-decision deconvolver of the `--hs` path (`k_hs_errors` + `k_hs_lock` + `k_hs_decode`: bytes, alignment and vote
phase equal the oracle's `dvb_deconvol_sync_hard`, votes every 32 chunks and every chunk, over two batches).  The same
binary built with -fsanitize=thread is the race check of these kernels (see test_ctl_kernels_cpu.py).  The GPU parity
tests check the same kernels through the C ABI; this is what can be said about them where there is no GPU.
"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"
CASES = ["rs", "rs_deint", "realign", "derand", "deconv", "sync", "hs"]


def _build(tmp, oracle_lib, tsan):
    src = open(os.path.join(ROOT, "leansdr_b200", "csrc", "k_fec.cu")).read()
    i = src.index("namespace {\n") + len("namespace {\n")
    j = src.index("}  // namespace\n")
    body = src[i:j].replace('#include "k_ctl_fec.cuh"', '#include "%s"' % os.path.join(ROOT, "leansdr_b200", "csrc", "k_ctl_fec.cuh"))
    assert "<<<" not in body and "k_rs" in body and "k_derand_out" in body
    inc = str(tmp / "k_fec_dev.inc")
    open(inc, "w").write(body)
    out = str(tmp / ("emu_fec_tsan" if tsan else "emu_fec"))
    cmd = ["g++", "-std=c++20", "-O1", "-pthread", "-w", "-I", CUDA_INC, '-DFEC_DEV_INC="%s"' % inc,
           os.path.join(ROOT, "tests", "emu", "emu_fec.cpp"), os.path.join(ROOT, "leansdr_b200", "csrc", "tables.cpp"),
           oracle_lib, "-Wl,-rpath," + os.path.dirname(oracle_lib), "-o", out]
    if tsan:
        cmd[1:1] = ["-g", "-fsanitize=thread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and tsan:
        pytest.skip("g++ cannot link ThreadSanitizer here: " + r.stderr[-300:])
    assert r.returncode == 0, r.stderr[-3000:]
    return out


@pytest.fixture(scope="module")
def oracle_lib(oracle):
    if shutil.which("g++") is None or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("g++ or the CUDA headers are not available")
    p = os.path.join(ROOT, "oracle", "liboracle.so")
    assert os.path.exists(p)
    return p


@pytest.mark.timeout(600)
def test_packet_kernels_equal_the_oracle_on_the_host(oracle_lib, tmp_path_factory):
    exe = _build(tmp_path_factory.mktemp("emu_fec"), oracle_lib, tsan=False)
    for case in CASES:
        for seed in (1, 2):
            r = subprocess.run([exe, case, str(seed)], capture_output=True, text=True, timeout=500)
            assert r.returncode == 0 and "equal" in r.stdout, f"{case} seed {seed}:\n{r.stderr[-2000:]}"


@pytest.mark.timeout(900)
def test_packet_kernels_are_race_free_under_thread_sanitizer(oracle_lib, tmp_path_factory):
    exe = _build(tmp_path_factory.mktemp("emu_fec_tsan"), oracle_lib, tsan=True)
    for case in CASES:
        if case == "realign":
            continue   # (k_realign and k_sync_flags share nothing between threads; `sync` runs them anyway)
        r = subprocess.run([exe, case, "4"], capture_output=True, text=True, timeout=800,
                           env=dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=0"))
        n = (r.stdout + r.stderr).count("WARNING: ThreadSanitizer")
        assert r.returncode == 0 and "equal" in r.stdout and n == 0, f"{case}: {n} reports\n{(r.stdout + r.stderr)[-3000:]}"


@pytest.mark.timeout(600)
def test_transmit_kernels_equal_the_oracle_on_the_host(oracle_lib, tmp_path_factory):
    """The kernels of the transmit chain (`leansdr_b200/csrc/tx.cu`: leantsgen packets, randomizer + rs_encoder,
    interleaver, dvb_convol for every code rate and symbol width; then, float for float, `k_tx_resample` =
    cstln_transmitter + fir_resampler + decimator at 6/5, 2/1 and 12/5 from symbols and from cf32 in two ragged
    launches, and `k_tx_amp2` / `k_tx_agc` / `k_tx_scale` = simple_agc in two pushes with the carried estimate, RRC
    taps and amplitude included) against the oracle's restatement of leandvbtx, which
    is pinned to the reference transmitter (tests/test_oracle_tx_cpu.py).  The device text is the first anonymous
    namespace of tx.cu; its one dynamic shared-memory declaration is pointed at the shim's buffer."""
    tmp = tmp_path_factory.mktemp("emu_tx")
    src = open(os.path.join(ROOT, "leansdr_b200", "csrc", "tx.cu")).read()
    i = src.index("namespace {\n") + len("namespace {\n")
    j = src.index("}  // namespace\n")
    body = src[i:j]
    assert "<<<" not in body and "k_tx_convol" in body and "extern __shared__ float2 s_taps[];" in body
    body = body.replace("extern __shared__ float2 s_taps[];", "float2 *s_taps = reinterpret_cast<float2 *>(emu::g_dyn_smem);")
    inc = str(tmp / "tx_dev.inc")
    open(inc, "w").write(body)
    exe = str(tmp / "emu_tx")
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-pthread", "-w", "-I", CUDA_INC, '-DTX_DEV_INC="%s"' % inc,
                        os.path.join(ROOT, "tests", "emu", "emu_tx.cpp"), os.path.join(ROOT, "leansdr_b200", "csrc", "tables.cpp"),
                        oracle_lib, "-Wl,-rpath," + os.path.dirname(oracle_lib), "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    for seed in (1, 2):
        r = subprocess.run([exe, str(seed)], capture_output=True, text=True, timeout=500)
        assert r.returncode == 0 and "equal" in r.stdout, f"seed {seed}:\n{r.stderr[-2000:]}"


@pytest.mark.timeout(600)
def test_telemetry_kernels_equal_the_oracle_on_the_host(oracle_lib, tmp_path_factory):
    """`k_spectrum.cu` on the host: k_meas_power (the reference's FFT butterfly order over shared memory, one barrier per
    stage, 4096 and 1024 points) and k_meas_ema (running average, band sums of do_cnr) against the oracle's cfft_engine
    and cnr_fft / spectrum, float for float (-ffp-contract=off; the shim's fmul / fadd are single IEEE operations like
    the device's _rn intrinsics): power bins, the average after every measurement, the CNR value.  Then the same run
    under ThreadSanitizer: the barrier-staged FFT is the kind of kernel a missing __syncthreads hides in."""
    tmp = tmp_path_factory.mktemp("emu_meas")
    src = open(os.path.join(ROOT, "leansdr_b200", "csrc", "k_spectrum.cu")).read()
    i = src.index("namespace {\n") + len("namespace {\n")
    j = src.index("}  // namespace\n")
    body = src[i:j]
    assert "<<<" not in body and "k_meas_power" in body and "k_meas_ema" in body
    inc = str(tmp / "meas_dev.inc")
    open(inc, "w").write(body)
    base = ["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-pthread", "-w", "-I", CUDA_INC, '-DMEAS_DEV_INC="%s"' % inc,
            os.path.join(ROOT, "tests", "emu", "emu_meas.cpp"), oracle_lib, "-Wl,-rpath," + os.path.dirname(oracle_lib)]
    exe = str(tmp / "emu_meas")
    r = subprocess.run(base + ["-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    for seed in (1, 2, 3):
        r = subprocess.run([exe, str(seed)], capture_output=True, text=True, timeout=500)
        assert r.returncode == 0 and "equal" in r.stdout, f"seed {seed}:\n{r.stderr[-2000:]}"
    tsan = str(tmp / "emu_meas_tsan")
    r = subprocess.run(base + ["-g", "-fsanitize=thread", "-o", tsan], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("g++ cannot link ThreadSanitizer here: " + r.stderr[-300:])
    r = subprocess.run([tsan, "4"], capture_output=True, text=True, timeout=500,
                       env=dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=0"))
    n = (r.stdout + r.stderr).count("WARNING: ThreadSanitizer")
    assert r.returncode == 0 and "equal" in r.stdout and n == 0, f"{n} reports\n{(r.stdout + r.stderr)[-3000:]}"


@pytest.mark.timeout(600)
def test_frontend_kernel_equals_the_oracle_on_the_host(oracle_lib, tmp_path_factory):
    """`k_frontend` (cconverter / scaler -> rotator -> fir_filter + decimator fused; the FIR stage of north_star) on
    the host against the oracle's chain of the same runnables, float for float, in six configurations: f32 with the 5
    real taps of the bench, u8 + rotator + 13 retuned (complex) taps + decimation 2, s16 through the decimator alone,
    scaled f32 with 31 retuned taps, u16 + rotator + 7 taps, s8 with 40 taps and decimation 5; two full tiles and a ragged
    one each.  The host-built rotator table and shifted taps (with the reference's unsigned tap-index quirk) must equal
    the oracle's first.  The bulk copy is done at issue time and the mbarrier is a release / acquire flag; then the same
    run under ThreadSanitizer."""
    tmp = tmp_path_factory.mktemp("emu_front")
    src = open(os.path.join(ROOT, "leansdr_b200", "csrc", "k_frontend.cu")).read()
    i = src.index("namespace {\n") + len("namespace {\n")
    j = src.index("}  // namespace\n")
    body = src[i:j]
    decl = "extern __shared__ __align__(128) unsigned char smem[];"
    assert "<<<" not in body and "k_frontend" in body and decl in body
    body = body.replace(decl, "unsigned char *smem = emu::g_dyn_smem;")
    inc = str(tmp / "front_dev.inc")
    open(inc, "w").write(body)
    base = ["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-pthread", "-w", "-I", CUDA_INC, '-DFRONT_DEV_INC="%s"' % inc,
            os.path.join(ROOT, "tests", "emu", "emu_front.cpp"), os.path.join(ROOT, "leansdr_b200", "csrc", "tables.cpp"),
            oracle_lib, "-Wl,-rpath," + os.path.dirname(oracle_lib)]
    exe = str(tmp / "emu_front")
    r = subprocess.run(base + ["-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    for seed in (1, 2, 3):
        r = subprocess.run([exe, str(seed)], capture_output=True, text=True, timeout=500)
        assert r.returncode == 0 and "equal" in r.stdout, f"seed {seed}:\n{r.stderr[-2000:]}"
    tsan = str(tmp / "emu_front_tsan")
    r = subprocess.run(base + ["-g", "-fsanitize=thread", "-o", tsan], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("g++ cannot link ThreadSanitizer here: " + r.stderr[-300:])
    r = subprocess.run([tsan, "4"], capture_output=True, text=True, timeout=500,
                       env=dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=0"))
    n = (r.stdout + r.stderr).count("WARNING: ThreadSanitizer")
    assert r.returncode == 0 and "equal" in r.stdout and n == 0, f"{n} reports\n{(r.stdout + r.stderr)[-3000:]}"


@pytest.mark.timeout(600)
def test_notch_detect_kernel_equals_the_oracle_on_the_host(oracle_lib, tmp_path_factory):
    """`k_notch_detect` (auto_notch::detect: cfft_engine's 4096-point FFT in shared memory, glibc's hypotf in double,
    the nslots largest bins with their neighbours blanked) on the host against the oracle's auto_notch made to detect on
    every block: 1..4 slots, cf32 and u8 input, tones on random bins (two of them adjacent), noise.  Then the same run
    under ThreadSanitizer."""
    tmp = tmp_path_factory.mktemp("emu_nd")
    src = open(os.path.join(ROOT, "leansdr_b200", "csrc", "k_notch.cu")).read()
    i = src.index("__device__ __forceinline__ float2 load_sample(")
    j = src.index("// ---------------------------------------------------------------------- apply")
    body = src[i:j]
    assert "<<<" not in body and "k_notch_detect" in body
    inc = str(tmp / "detect_dev.inc")
    open(inc, "w").write(body)
    base = ["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-pthread", "-w", "-I", CUDA_INC, '-DDETECT_DEV_INC="%s"' % inc,
            os.path.join(ROOT, "tests", "emu", "emu_notch_detect.cpp"), oracle_lib, "-Wl,-rpath," + os.path.dirname(oracle_lib)]
    exe = str(tmp / "emu_nd")
    r = subprocess.run(base + ["-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe, "1"], capture_output=True, text=True, timeout=500)
    assert r.returncode == 0 and "equal" in r.stdout, r.stderr[-2000:]
    tsan = str(tmp / "emu_nd_tsan")
    r = subprocess.run(base + ["-g", "-fsanitize=thread", "-o", tsan], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("g++ cannot link ThreadSanitizer here: " + r.stderr[-300:])
    r = subprocess.run([tsan, "2"], capture_output=True, text=True, timeout=500,
                       env=dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=0"))
    n = (r.stdout + r.stderr).count("WARNING: ThreadSanitizer")
    assert r.returncode == 0 and "equal" in r.stdout and n == 0, f"{n} reports\n{(r.stdout + r.stderr)[-3000:]}"


@pytest.mark.timeout(900)
def test_notch_kernels_equal_the_oracle_on_the_host(oracle_lib, tmp_path_factory):
    """`k_notch_guess` + `k_notch_apply` + `k_notch_verify` (k_notch.cu) on the host against the oracle's
    auto_notch::process, float for float, with 1, 2 and 3 slots: one exact segment over the batch; one segment per block
    with guessed start states and two warm-up blocks (every segment merges bit for bit); and without warm-up blocks,
    where nothing merges and every segment is re-run from its predecessor's exit state in rounds (the scheme of
    run_notch / notch_verify_repair) -- the streams and the carried estimates are the oracle's either way.  The
    asynchronous row copies are done at issue time.  Then the same run under ThreadSanitizer."""
    tmp = tmp_path_factory.mktemp("emu_na")
    csrc = os.path.join(ROOT, "leansdr_b200", "csrc")
    src = open(os.path.join(csrc, "k_notch.cu")).read()
    common = open(os.path.join(csrc, "notch_common.cuh")).read()
    ci = common.index("namespace {\n") + len("namespace {\n")
    cj = common.index("}  // namespace\n}  // namespace ldvb")
    i = src.index("namespace {\n") + len("namespace {\n")
    j = src.index("template <int FMT, int NSLOTS>\ncudaError_t launch_apply_t(")
    vi = src.index("// entry(j) == exit(j-1), bit for bit, for every segment that started from a guess.")
    vj = src.index("}  // namespace\n", vi)
    body, verify = src[i:j], src[vi:vj]
    decl = "extern __shared__ __align__(128) unsigned char smem[];"
    assert decl in body and "<<<" not in body and "<<<" not in verify and "k_notch_verify" in verify
    inc = str(tmp / "notch_dev.inc")
    open(inc, "w").write(common[ci:cj] + "\n" + body.replace(decl, "unsigned char *smem = emu::g_dyn_smem;") + "\n" + verify)
    base = ["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-pthread", "-w", "-I", CUDA_INC, '-DNOTCH_DEV_INC="%s"' % inc,
            os.path.join(ROOT, "tests", "emu", "emu_notch_apply.cpp"), oracle_lib, "-Wl,-rpath," + os.path.dirname(oracle_lib)]
    exe = str(tmp / "emu_na")
    r = subprocess.run(base + ["-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe, "1"], capture_output=True, text=True, timeout=800)
    assert r.returncode == 0 and "equal" in r.stdout, r.stderr[-2000:]
    assert "0 warm-up blocks: 8 of 9 segments repaired" in r.stderr      # the repair path ran
    tsan = str(tmp / "emu_na_tsan")
    r = subprocess.run(base + ["-g", "-fsanitize=thread", "-o", tsan], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("g++ cannot link ThreadSanitizer here: " + r.stderr[-300:])
    r = subprocess.run([tsan, "2", "quick"], capture_output=True, text=True, timeout=800,
                       env=dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=0"))
    n = (r.stdout + r.stderr).count("WARNING: ThreadSanitizer")
    assert r.returncode == 0 and "equal" in r.stdout and n == 0, f"{n} reports\n{(r.stdout + r.stderr)[-3000:]}"


@pytest.mark.timeout(900)
def test_fused_notch_fir_kernel_equals_the_oracle_on_the_host(oracle_lib, tmp_path_factory):
    """The DEFAULT notch kernel of the receive chain, `k_notch_fir` + `k_fir_edges` (k_notchfir.cu: auto_notch and the
    fir_filter behind it fused; a chain warp and eight worker warps per 16 segments, a block barrier per 64-sample tile,
    five stages of asynchronous row copies), behind `k_notch_guess` and in front of `k_notch_verify`, on the host against
    the oracle's auto_notch -> fir_filter, float for float: two consecutive batches (carried notched samples and
    estimates), segments that merge after two warm-up blocks and segments without warm-up that are all re-run from their
    predecessors' exit states (rewriting their edge samples), 5 real taps (the bench configuration), 13 retuned taps,
    plain notch through the same kernel, 1 and 2 slots, the telemetry dump.  Then under ThreadSanitizer."""
    tmp = tmp_path_factory.mktemp("emu_nf")
    csrc = os.path.join(ROOT, "leansdr_b200", "csrc")
    common = open(os.path.join(csrc, "notch_common.cuh")).read()
    ci = common.index("namespace {\n") + len("namespace {\n")
    cj = common.index("}  // namespace\n}  // namespace ldvb")
    decl = "extern __shared__ __align__(128) unsigned char smem[];"
    src = open(os.path.join(csrc, "k_notch.cu")).read()
    i = src.index("namespace {\n") + len("namespace {\n")
    j = src.index("template <int FMT, int NSLOTS>\ncudaError_t launch_apply_t(")
    vi = src.index("// entry(j) == exit(j-1), bit for bit, for every segment that started from a guess.")
    vj = src.index("}  // namespace\n", vi)
    inc1 = str(tmp / "notch_dev.inc")
    open(inc1, "w").write(common[ci:cj] + "\n" + src[i:j].replace(decl, "unsigned char *smem = emu::g_dyn_smem;") + "\n" + src[vi:vj])
    srcf = open(os.path.join(csrc, "k_notchfir.cu")).read()
    i = srcf.index("namespace {\n") + len("namespace {\n")
    j = srcf.index("template <int FMT, int NSLOTS, bool FIR>\ncudaError_t launch_nf_t(")
    body = srcf[i:j]
    assert decl in body and "<<<" not in body and "k_notch_fir" in body and "k_fir_edges" in body
    inc2 = str(tmp / "notchfir_dev.inc")
    open(inc2, "w").write(common[ci:cj] + "\n" + body.replace(decl, "unsigned char *smem = emu::g_dyn_smem;"))
    base = ["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-pthread", "-w", "-I", CUDA_INC, '-DNOTCH_DEV_INC="%s"' % inc1,
            '-DNOTCHFIR_DEV_INC="%s"' % inc2, os.path.join(ROOT, "tests", "emu", "emu_notchfir.cpp"),
            os.path.join(csrc, "tables.cpp"), oracle_lib, "-Wl,-rpath," + os.path.dirname(oracle_lib)]
    exe = str(tmp / "emu_nf")
    r = subprocess.run(base + ["-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe, "1"], capture_output=True, text=True, timeout=800)
    assert r.returncode == 0 and "equal" in r.stdout, r.stderr[-2000:]
    assert "0 warm-up blocks: 24 segments repaired" in r.stderr          # the repair path ran
    tsan = str(tmp / "emu_nf_tsan")
    r = subprocess.run(base + ["-g", "-fsanitize=thread", "-o", tsan], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("g++ cannot link ThreadSanitizer here: " + r.stderr[-300:])
    r = subprocess.run([tsan, "2", "quick"], capture_output=True, text=True, timeout=800,
                       env=dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=0"))
    n = (r.stdout + r.stderr).count("WARNING: ThreadSanitizer")
    assert r.returncode == 0 and "equal" in r.stdout and n == 0, f"{n} reports\n{(r.stdout + r.stderr)[-3000:]}"


def _rx_device_text(tmp):
    """The device code of k_rx.cu (its first anonymous namespace: every kernel of the receiver stage, k_ctl_rx.cuh
    included) with the three inline-PTX statements replaced by what they mean and the dynamic shared memory pointed at
    the shim's buffer."""
    csrc = os.path.join(ROOT, "leansdr_b200", "csrc")
    src = open(os.path.join(csrc, "k_rx.cu")).read()
    i = src.index("namespace {\n") + len("namespace {\n")
    j = src.index("}  // namespace\n\ncudaError_t launch_rx_power(")
    body = src[i:j]
    subs = [
        ("extern __shared__ __align__(128) unsigned char smem[];", "unsigned char *smem = emu::g_dyn_smem;", 2),
        ('asm volatile("ld.shared.s16 %0, [%1];" : "=h"(v) : "r"(addr));',
         "v = *reinterpret_cast<const short *>(emu::g_dyn_smem + addr);", 1),
        ('asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(Ii) : "f"(I));', "Ii = f2i_trunc(I);", 1),
        ('asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(Qi) : "f"(Q));', "Qi = f2i_trunc(Q);", 1),
        ('#include "k_ctl_rx.cuh"', '#include "%s"' % os.path.join(csrc, "k_ctl_rx.cuh"), 1),
    ]
    for old, new, count in subs:
        assert body.count(old) == count, (old, body.count(old))
        body = body.replace(old, new)
    # what is left of inline PTX is the evict_last table load, compiled out by LDVB_RX_TRIG_EVICT_LAST=0 (then __ldg)
    assert body.count("asm") == 1 and "<<<" not in body
    for k in ("k_rx_serial", "k_rx(", "k_rx_stitch(", "k_rx_compact(", "rx_warm_state"):
        assert k in body, k
    inc = str(tmp / "rx_dev.inc")
    open(inc, "w").write(body)
    return inc


@pytest.mark.timeout(900)
def test_receiver_kernels_equal_the_oracle_on_the_host(oracle_lib, tmp_path_factory):
    """The receiver stage (k_rx.cu: the dominant kernel of the chain) on the host against the oracle's cstln_receiver,
    field for field: `k_rx_serial` (EXACT mode: softsymbols, end state, sampled-point tap, measurement rows), `k_rx`
    with one lane per time span (span 0 from the true state equals the oracle at once; every other span is then re-run
    from its predecessor's end state, the repair path of FAST mode, and the spans with their verification overlaps
    are the oracle's stream), `k_rx_stitch` on those exact spans (every seam verifies under the strict rule with zero
    state difference), `k_rx_stitch_pair` (the seam between two ranks: same judgement), `k_rx_power`, `k_rx_plan_local/_apply` + `k_rx_compact` (the contiguous stream is the oracle's).  QPSK with the
    arithmetic slicer and the phase-error column in shared memory (the bench configuration), the cell-table slicer,
    nearest / linear / RRC samplers, 8PSK and 16APSK, and the integer receiver of `--hs` (fast_qpsk_receiver: hard symbols,
    loop state, frequency rows).  FAST mode on the kernels themselves, scheduled as run_receiver does (serial settling
    pass, 64 speculative spans from the carried frequency / AGC after 4 warm-up chunks, strict seams, exact re-run behind a
    failed seam, plan with the spans' rotations, compaction): the stitched stream carries the oracle's hard decisions --
    all of them on a waveform with open eyes, all but marginal ones (both costs under 600 of ~11000) under heavy
    inter-symbol interference -- while the soft costs differ with the AGC state a span started from (printed).  Then the
    same kernels under ThreadSanitizer."""
    tmp = tmp_path_factory.mktemp("emu_rx")
    inc = _rx_device_text(tmp)
    csrc = os.path.join(ROOT, "leansdr_b200", "csrc")
    base = ["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-pthread", "-w", "-I", CUDA_INC, '-DRX_DEV_INC="%s"' % inc,
            "-DLDVB_RX_TRIG_EVICT_LAST=0", os.path.join(ROOT, "tests", "emu", "emu_rx.cpp"),
            os.path.join(csrc, "tables.cpp"), oracle_lib, "-Wl,-rpath," + os.path.dirname(oracle_lib)]
    exe = str(tmp / "emu_rx")
    r = subprocess.run(base + ["-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe, "1"], capture_output=True, text=True, timeout=800)
    assert r.returncode == 0 and "equal" in r.stdout, r.stderr[-3000:]
    assert r.stderr.count("equal so far: yes") == 9
    assert "FAST mode" in r.stderr and "hard decisions differing from the oracle: 0," in r.stderr
    tsan = str(tmp / "emu_rx_tsan")
    r = subprocess.run(base + ["-g", "-fsanitize=thread", "-o", tsan], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("g++ cannot link ThreadSanitizer here: " + r.stderr[-300:])
    r = subprocess.run([tsan, "3", "quick"], capture_output=True, text=True, timeout=800,
                       env=dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=0"))
    n = (r.stdout + r.stderr).count("WARNING: ThreadSanitizer")
    assert r.returncode == 0 and "equal" in r.stdout and n == 0, f"{n} reports\n{(r.stdout + r.stderr)[-3000:]}"
