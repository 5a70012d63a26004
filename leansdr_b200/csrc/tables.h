// tables.h -- host-side constant builders for the B200 DVB-S receive path.
//
// Everything that depends on glibc libm (cosf/sinf/atan2f/sqrt) is computed
// ONCE on the host with the same expressions, types and evaluation order as
// the reference uses, then uploaded; the device never recomputes a
// transcendental that feeds the bit-exact path.  Citations: file:line under
// /root/reference/src/leansdr/ unless stated otherwise.
#pragma once
#include <cstdint>
#include <vector>

namespace ldvb {

struct CstlnCell { int16_t cost, symbol, phase_error, pad; };  // 8 bytes

struct Cstln {
  std::vector<CstlnCell> cells;       // 65536, index = (u8)I*256 + (u8)Q  (sdr.h:486)
  std::vector<int8_t> sym_re, sym_im; // constellation points (sdr.h:313-339)
  int nsymbols = 0, nrotations = 0;
  // rot[k][s]: symbol index of point s rotated by k*360/nrotations degrees.
  std::vector<std::vector<uint8_t>> rot;
};

// kind = LDVB_CSTLN_*, fec = LDVB_FEC* (APSK ring ratios depend on it, dvb.h:45-81).
Cstln make_cstln(int kind, int fec, bool harden);        // sdr.h:313-573
std::vector<float> make_trig16();                        // math.h:95-111, 65536 x {cos,sin}
void make_rs_tables(uint8_t exp512[512], uint8_t log256[256]);  // rs.h:49-60
std::vector<uint8_t> make_derand_pattern();              // dvb.h:1116-1129

// leandvb.cc:353-378: returns normalised low-pass taps and the decimation.
std::vector<float> design_resampler(float Fs, float Fm, float rolloff, float rej,
                                    unsigned decim_opt, int *decim_out);
// dsp.h:270-280 (including the unsigned tap-index quirk).
std::vector<float> shift_taps(const std::vector<float> &coeffs, float freq);
// leandvb.cc:437-456 + filtergen.h:68-92.
std::vector<float> design_rrc(float Fs, float Fm, float rolloff, float rej,
                              int steps_opt, int *steps_out);
// leandvbtx.cc:131-138: RRC interpolation taps scaled by normalize_power(amp / cstln_amp).
std::vector<float> design_tx_rrc(int interp, float rolloff, float rrc_rej, float amp);
// fir_resampler::set_freq (dsp.h:352-361): complex taps {c*cosf(a), c*sinf(a)}, a = 2*pi*f*i.
std::vector<float> shift_taps_resampler(const std::vector<float> &coeffs, float freq);
// sdr.h:1231-1241: 65536 x cos then 65536 x sin.
std::vector<float> make_rotator_lut(float freq);

// fast_qpsk_receiver::init_lookup_tables (sdr.h:1144-1164), index = (u8)re * 256 + (u8)im:
//   polar[i] = angle (u16) | magnitude (u8) << 16; rect[angle8 * 256 + r] = re | im << 8;
//   sincos[angle16] = re | im << 8.
struct HsTables { std::vector<uint32_t> polar; std::vector<uint16_t> rect, sincos; };
HsTables make_hs_tables();

struct DeconvPolys {
  int punctperiod = 0, punctweight = 0;
  uint64_t deconv[8] = {0}, deconv2[8] = {0};
  uint8_t hyp_lut[4][4] = {{0}};   // [hypothesis][symbol&3] -> 2 IQ bits (dvb.h:309-360)
};
bool make_deconv(int fec, DeconvPolys *out);             // dvb.h:124-292

struct Trellis {
  int bits_in = 0, bits_out = 0, nus = 0, ncs = 0;
  int path_nbits = 0, path_depth = 0; bool path32 = false;
  std::vector<uint8_t> pred, us;   // [64*ncs]; pred==65 means no branch (viterbi.h:46,52-55)
};
bool make_trellis(int fec, Trellis *out);                // viterbi.h:61-92, dvb.h:1180-1212

struct VitSyncs {
  int nsyncs = 0, nshifts = 0, bps = 0;
  std::vector<int> shift;                  // [nsyncs]
  std::vector<std::vector<uint8_t>> map;   // [nsyncs][nsymbols]
};
VitSyncs make_vitsyncs(const Cstln &c, const Trellis &t); // dvb.h:1236-1297, 1336-1351

}  // namespace ldvb
