"""Regenerates tests/golden/* from the UNMODIFIED reference (run in the dev container,
where /root/reference is mounted and `make -C oracle ref` has built oracle/_ref).

  c1_160.u8            leantsgen -c 160 | leandvbtx -f 6/5 --power 37.5 --agc | leanchansim --ou8
  c1_160.ts            leandvb --u8 -f 2400e3 --sr 2000e3 --cr 1/2          < c1_160.u8
  c1_160_resample.ts   leandvb --u8 ... --resample                           < c1_160.u8
  c1_160_anf0.ts       leandvb --u8 ... --anf 0                              < c1_160.u8
  c1_160_hs.ts         leandvb --u8 ... --hs                                 < c1_160.u8
                       (`python make_golden.py hs` regenerates only this file)
  c1_160_taps.json     sha256 + sizes of every stream tapped by oracle/_ref/ref_tap (default flags)
  tables.json          sha256 of the constant tables dumped by oracle/_ref/ref_tables
                       (`python make_golden.py tables` regenerates only this file and the small verbatim tables)
  c1_160_spectrum_fs100k.f32  spectrum rows (p_spectrum) of ref_tap --u8 -f 100000 --sr 83333 < c1_160.u8
  kat.json             known answers quoted in SURVEY.md 8(c)
  tx_kat.json          sha256 + length of oracle/_ref/leandvbtx's cf32 output for tests/tx_cases.py
                       (`python make_golden.py tx` regenerates only this file)
"""
import hashlib, json, os, subprocess, sys, tempfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O
from tests import vectors as V

def sha(b): return hashlib.sha256(b).hexdigest()

def spectrum_golden(iq):
    """spectrum<f32> (sdr.h:1347-1404): at -f 100000 the 291 k-sample fixture yields 2 rows."""
    d = tempfile.mkdtemp()
    subprocess.run([O.ref_bin("ref_tap"), "--u8", "-f", "100000", "--sr", "83333", "--tap-dir", d], input=iq.tobytes(),
                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True)
    rows = np.fromfile(os.path.join(d, "spectrum.f32"), np.float32)
    assert rows.size == 2 * 1024
    rows.tofile(os.path.join(HERE, "c1_160_spectrum_fs100k.f32"))

def tx_golden():
    from tests.tx_cases import TX_CASES
    kat = {}
    for name, npk, cst, cr, ratio, power, agc, rolloff in TX_CASES:
        args = [O.ref_bin("leandvbtx"), "--const", cst, "--cr", cr, "-f", ratio, "--power", power, "--roll-off", str(rolloff)]
        if agc: args.append("--agc")
        out = subprocess.run(args, input=V.ts_packets(npk).tobytes(), stdout=subprocess.PIPE, check=True).stdout
        kat[name] = {"samples": len(out) // 8, "sha256": sha(out), "cmd": " ".join(["leantsgen -c %d |" % npk, "leandvbtx"] + args[1:])}
    json.dump(kat, open(os.path.join(HERE, "tx_kat.json"), "w"), indent=1)

def hs_golden():
    iq = np.fromfile(os.path.join(HERE, "c1_160.u8"), np.uint8)
    V.ref_leandvb(iq, ["--u8", "-f", "2400e3", "--sr", "2000e3", "--cr", "1/2", "--hs"]).tofile(os.path.join(HERE, "c1_160_hs.ts"))

def tables_golden():
    d2 = tempfile.mkdtemp()
    subprocess.run([O.ref_bin("ref_tables"), d2], check=True)
    tabs = {f: {"bytes": os.path.getsize(os.path.join(d2, f)), "sha256": sha(open(os.path.join(d2, f), "rb").read())}
            for f in sorted(os.listdir(d2))}
    json.dump(tabs, open(os.path.join(HERE, "tables.json"), "w"), indent=1)
    # Small tables are committed verbatim (libm independent).
    for f in ("rs_exp.u8", "rs_log.u8", "rs_gen.u8", "derand.u8", "deconv_12.u64", "deconv_34.u64", "deconv_78.u64",
              "trellis_12.bin", "vitmap_qpsk12.u8", "cstln_qpsk_symbols.s8", "cstln_16apsk34_symbols.s8",
              "cstln_64apske_symbols.s8", "cstln_256qam_symbols.s8", "vitmap_16apsk34.u8"):
        open(os.path.join(HERE, f), "wb").write(open(os.path.join(d2, f), "rb").read())

def main():
    if sys.argv[1:] == ["tables"]:
        tables_golden()
        print("tables golden regenerated")
        return
    if sys.argv[1:] == ["hs"]:
        hs_golden()
        print("hs golden regenerated")
        return
    if sys.argv[1:] == ["tx"]:
        tx_golden()
        print("tx golden regenerated")
        return
    tx_golden()
    iq = V.ref_iq(160, fmt="u8")
    iq.tofile(os.path.join(HERE, "c1_160.u8"))
    base = ["--u8", "-f", "2400e3", "--sr", "2000e3", "--cr", "1/2"]
    for name, extra in (("c1_160.ts", []), ("c1_160_resample.ts", ["--resample"]), ("c1_160_anf0.ts", ["--anf", "0"])):
        V.ref_leandvb(iq, base + extra).tofile(os.path.join(HERE, name))
    d = tempfile.mkdtemp()
    subprocess.run([O.ref_bin("ref_tap"), *base, "--tap-dir", d], input=iq.tobytes(), stdout=subprocess.PIPE, check=True)
    taps = {}
    for f in sorted(os.listdir(d)):
        if f.endswith(".txt"): continue
        b = open(os.path.join(d, f), "rb").read()
        taps[f] = {"bytes": len(b), "sha256": sha(b)}
    json.dump(taps, open(os.path.join(HERE, "c1_160_taps.json"), "w"), indent=1)
    spectrum_golden(iq)
    hs_golden()
    tables_golden()
    kat = {"deconv_fec12": "0x3ba", "rs_gen": "01 3b 0d 68 bd 44 d1 1e 08 a3 41 29 e5 62 32 24 3b",
           "qpsk_points": [[53, 53], [53, -53], [-53, 53], [-53, -53]],
           "lookup_53_53": [-11236, 0, 0], "lookup_10_m3": [-636, 1, 5151]}
    json.dump(kat, open(os.path.join(HERE, "kat.json"), "w"), indent=1)
    print("golden regenerated")

if __name__ == "__main__":
    main()
