"""Transmit-chain test cases shared by the CPU (oracle vs reference) and GPU (product vs oracle) tests:
(name, packets, --const, --cr, -f, --power, --agc, --roll-off)."""
TX_CASES = [
    ("qpsk12_6_5_agc", 200, "QPSK", "1/2", "6/5", "37.5", True, 0.35),     # the bench waveform (C1/C2)
    ("qpsk12_2", 200, "QPSK", "1/2", "2", "0", False, 0.35),               # leandvbtx defaults
    ("qpsk78_2_agc", 200, "QPSK", "7/8", "2", "37.5", True, 0.35),         # C3 (broadcast rate)
    ("8psk23_4_agc", 150, "8PSK", "2/3", "4", "10", True, 0.35),
    ("qpsk23_3_2_agc", 150, "QPSK", "2/3", "3/2", "37.5", True, 0.35),     # 2/3 on QPSK runs as 4/6
    ("bpsk12_2_agc", 100, "BPSK", "1/2", "2", "37.5", True, 0.35),
    ("qpsk34_5_agc", 100, "QPSK", "3/4", "5", "37.5", True, 0.2),
    ("qpsk56_5_3", 100, "QPSK", "5/6", "5/3", "37.5", False, 0.2),
    # the rest of cstln_lut<256>::predef (sdr.h:305-311), code rates the block length allows (dvb.h:582-584)
    ("16apsk34_2_agc", 100, "16APSK", "3/4", "2", "37.5", True, 0.35),
    ("64apske23_3_agc", 100, "64APSKe", "2/3", "3", "37.5", True, 0.35),   # 2/3 on 64 points runs as 4/6
    ("16qam34_2_agc", 100, "16QAM", "3/4", "2", "37.5", True, 0.35),
    ("64qam56_2_agc", 100, "64QAM", "5/6", "2", "37.5", True, 0.35),
    ("256qam78_2_agc", 100, "256QAM", "7/8", "2", "37.5", True, 0.35),
]
# "Code rate not suitable for this constellation" (dvb.h:582-584): coded bits per block not a multiple of bits per symbol.
TX_REJECTED = [("16APSK", "2/3"), ("32APSK", "5/6"), ("8PSK", "1/2"), ("16APSK", "1/2")]
