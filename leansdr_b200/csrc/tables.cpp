// tables.cpp -- host-side constant builders (see tables.h).
//
// Compiled with -ffp-contract=off: the values must round exactly like the
// reference's baseline x86-64 build (scalar SSE2, no FMA).
#include "tables.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace ldvb {

namespace {

constexpr float kCstlnAmp = 75.0f;  // sdr.h:297

inline unsigned par64(uint64_t v) { return (unsigned)__builtin_parityll(v); }

inline int bitlen(uint64_t v) { return v ? 64 - __builtin_clzll(v) : 0; }

// One constellation point at angle i/n of a turn, radius r (sdr.h:492-495):
// the angle is formed as float*int, widened by the double constant M_PI,
// narrowed to float; the products r*cosf(a)*75 are float and truncate to s8.
void put_point(Cstln &c, int s, float r, int n, float i) {
  float a = (float)((double)(i * 2) * M_PI / n);
  c.sym_re[s] = (int8_t)(r * cosf(a) * kCstlnAmp);
  c.sym_im[s] = (int8_t)(r * sinf(a) * kCstlnAmp);
}

// polar2() (sdr.h:497-504): four points at angles a*pi (float angle, double M_PI).
void put_polar2(Cstln &c, int i, float r, float a0, float a1, float a2, float a3) {
  const float a[4] = {a0, a1, a2, a3};
  for (int j = 0; j < 4; ++j) {
    float phi = (float)((double)a[j] * M_PI);
    c.sym_re[i + j] = (int8_t)(r * cosf(phi) * kCstlnAmp);
    c.sym_im[i + j] = (int8_t)(r * sinf(phi) * kCstlnAmp);
  }
}

// make_qam() (sdr.h:505-528): m x m grid, x-major, unit average power.
void put_qam(Cstln &c, int n) {
  c.nsymbols = n; c.nrotations = 4;
  int m = (int)sqrtl((long double)n);
  float scale;
  {
    int q = m / 2;
    float avgpower = (float)(2 * (q * 0.25 + (q - 1) * q / 2 + (q - 1) * q * (2 * q - 1) / 6) / q);
    scale = (float)(1.0 / sqrtf(avgpower));
  }
  int s = 0;
  for (int x = 0; x < m; ++x)
    for (int y = 0; y < m; ++y) {
      float I = x - (float)(m - 1) / 2;
      float Q = y - (float)(m - 1) / 2;
      c.sym_re[s] = (int8_t)(I * scale * kCstlnAmp);
      c.sym_im[s] = (int8_t)(Q * scale * kCstlnAmp);
      ++s;
    }
}

// make_dvbs2_constellation (dvb.h:45-81): APSK ring ratios per code rate.  Only the DVB-S
// rates exist behind this ABI; a rate the reference rejects returns false.
bool apsk_gammas(int kind, int fec, float *g1, float *g2, float *g3) {
  *g1 = *g2 = *g3 = 1;
  if (kind == 3) {         // APSK16, EN 302 307 table 9
    switch (fec) {
      case 1: case 2: *g1 = 3.15; return true;   // 2/3, 4/6
      case 3: *g1 = 2.85; return true;           // 3/4
      case 4: *g1 = 2.70; return true;           // 5/6
      default: return false;
    }
  }
  if (kind == 4) {         // APSK32, table 10
    switch (fec) {
      case 3: *g1 = 2.84; *g2 = 5.27; return true;
      case 4: *g1 = 2.64; *g2 = 4.64; return true;
      default: return false;
    }
  }
  if (kind == 5) { *g1 = 2.4; *g2 = 4.3; *g3 = 7; }   // APSK64E, EN 302 307-2 table 13f
  return true;
}

}  // namespace

Cstln make_cstln(int kind, int fec, bool harden) {
  Cstln c;
  c.sym_re.assign(256, 0);
  c.sym_im.assign(256, 0);
  float g1, g2, g3;
  if (!apsk_gammas(kind, fec, &g1, &g2, &g3)) return c;
  switch (kind) {
    case 0:  // BPSK at 45 degrees (sdr.h:315-327)
      c.nsymbols = 2; c.nrotations = 2;
      put_point(c, 0, 1, 8, 1); put_point(c, 1, 1, 8, 5);
      break;
    case 1: {  // QPSK (sdr.h:328-339)
      static const float q[4] = {0.5f, 3.5f, 1.5f, 2.5f};
      c.nsymbols = 4; c.nrotations = 4;
      for (int s = 0; s < 4; ++s) put_point(c, s, 1, 4, q[s]);
      break;
    }
    case 2: {  // 8PSK (sdr.h:340-354)
      static const int o[8] = {1, 0, 4, 5, 2, 7, 3, 6};
      c.nsymbols = 8; c.nrotations = 8;
      for (int s = 0; s < 8; ++s) put_point(c, s, 1, 8, (float)o[s]);
      break;
    }
    case 3: {  // 16APSK (sdr.h:355-381): 12 points on the outer ring, 4 inside
      static const float o[12] = {1.5f, 10.5f, 4.5f, 7.5f, 0.5f, 11.5f, 5.5f, 6.5f, 2.5f, 9.5f, 3.5f, 8.5f};
      static const float q[4] = {0.5f, 3.5f, 1.5f, 2.5f};
      float r1 = sqrtf(4 / (1 + 3 * g1 * g1));
      float r2 = g1 * r1;
      c.nsymbols = 16; c.nrotations = 4;
      for (int s = 0; s < 12; ++s) put_point(c, s, r2, 12, o[s]);
      for (int s = 0; s < 4; ++s) put_point(c, 12 + s, r1, 4, q[s]);
      break;
    }
    case 4: {  // 32APSK (sdr.h:382-424): rings of 4, 12 and 16 points
      // {ring (1..3), position}; ring 1 has 4 positions, ring 2 twelve, ring 3 sixteen.
      static const struct { int ring; float i; } t[32] = {
        {2, 1.5f}, {2, 2.5f}, {2, 10.5f}, {2, 9.5f}, {2, 4.5f}, {2, 3.5f}, {2, 7.5f}, {2, 8.5f},
        {3, 1}, {3, 3}, {3, 14}, {3, 12}, {3, 6}, {3, 4}, {3, 9}, {3, 11},
        {2, 0.5f}, {1, 0.5f}, {2, 11.5f}, {1, 3.5f}, {2, 5.5f}, {1, 1.5f}, {2, 6.5f}, {1, 2.5f},
        {3, 0}, {3, 2}, {3, 15}, {3, 13}, {3, 7}, {3, 5}, {3, 8}, {3, 10}};
      float r1 = sqrtf(8 / (1 + 3 * g1 * g1 + 4 * g2 * g2));
      float r2 = g1 * r1;
      float r3 = g2 * r1;
      c.nsymbols = 32; c.nrotations = 4;
      for (int s = 0; s < 32; ++s) {
        if (t[s].ring == 1) put_point(c, s, r1, 4, t[s].i);
        else if (t[s].ring == 2) put_point(c, s, r2, 12, t[s].i);
        else put_point(c, s, r3, 16, t[s].i);
      }
      break;
    }
    case 5: {  // 64APSK "E" (sdr.h:425-452): rings of 4, 12, 20 and 28 points
      float r1 = sqrtf(64 / (4 + 12 * g1 * g1 + 20 * g2 * g2 + 28 * g3 * g3));
      float r2 = g1 * r1;
      float r3 = g2 * r1;
      float r4 = g3 * r1;
      c.nsymbols = 64; c.nrotations = 4;
      // Groups of four points that are mirror images of each other: the first angle (in
      // units of pi/den) fixes the other three (2 - a, 1 - a, 1 + a).
      static const struct { int ring; int num, den; } t[16] = {
        {4, 1, 4}, {4, 13, 28}, {4, 1, 28}, {1, 1, 4}, {4, 9, 28}, {4, 11, 28}, {3, 1, 20}, {2, 1, 12},
        {4, 5, 28}, {3, 9, 20}, {4, 3, 28}, {2, 5, 12}, {3, 1, 4}, {3, 7, 20}, {3, 3, 20}, {2, 1, 4}};
      const float rr[5] = {0, r1, r2, r3, r4};
      for (int g = 0; g < 16; ++g) {
        const double n = t[g].num, d = t[g].den;
        put_polar2(c, 4 * g, rr[t[g].ring], (float)(n / d), (float)((2 * d - n) / d),
                   (float)((d - n) / d), (float)((d + n) / d));
      }
      break;
    }
    case 6: put_qam(c, 16); break;
    case 7: put_qam(c, 64); break;
    case 8: put_qam(c, 256); break;
    default:
      return c;  // nsymbols == 0 signals "unsupported"
  }
  // Decision table over the 256x256 integer grid (sdr.h:526-561).
  c.cells.resize(65536);
  for (int I = -128; I < 128; ++I) {
    for (int Q = -128; Q < 128; ++Q) {
      int32_t d_best = 2 * 256 * 256, d_second = 2 * 256 * 256;
      int best = 0;
      for (int s = 0; s < c.nsymbols; ++s) {
        int32_t di = I - c.sym_re[s], dq = Q - c.sym_im[s];
        int32_t d2 = di * di + dq * dq;
        if (d2 < d_best) { d_second = d_best; d_best = d2; best = s; }
        else if (d2 < d_second) d_second = d2;
      }
      if (d_best > 32767) d_best = 32767;
      if (d_second > 32767) d_second = 32767;
      CstlnCell &cell = c.cells[(size_t)(uint8_t)I * 256 + (uint8_t)Q];
      cell.cost = (int16_t)(d_best - d_second);
      cell.symbol = (int16_t)best;
      float ph_sym = atan2f((float)c.sym_im[best], (float)c.sym_re[best]);
      float ph_err = atan2f((float)Q, (float)I) - ph_sym;
      // (s32) is a 64-bit signed long in the reference; stored modulo 2^16.
      long long wide = (long long)((double)(ph_err * 65536) / (2 * M_PI));
      cell.phase_error = (int16_t)(uint16_t)(unsigned long long)wide;
      cell.pad = 0;
    }
  }
  if (harden)  // sdr.h:564-571
    for (auto &cell : c.cells) {
      if (cell.cost < 0) cell.cost = -1;
      if (cell.cost > 0) cell.cost = 1;
    }
  // Rotation permutations (used to resolve the phase ambiguity between
  // concurrently demodulated time spans): nearest symbol of each rotated point.
  c.rot.assign(c.nrotations, std::vector<uint8_t>(c.nsymbols, 0));
  for (int k = 0; k < c.nrotations; ++k) {
    double ang = 2 * M_PI * k / c.nrotations;
    for (int s = 0; s < c.nsymbols; ++s) {
      double x = c.sym_re[s] * cos(ang) - c.sym_im[s] * sin(ang);
      double y = c.sym_re[s] * sin(ang) + c.sym_im[s] * cos(ang);
      int best = 0; double bd = 1e30;
      for (int t = 0; t < c.nsymbols; ++t) {
        double d = (x - c.sym_re[t]) * (x - c.sym_re[t]) + (y - c.sym_im[t]) * (y - c.sym_im[t]);
        if (d < bd) { bd = d; best = t; }
      }
      c.rot[k][s] = (uint8_t)best;
    }
  }
  return c;
}

std::vector<float> make_trig16() {
  std::vector<float> t(2 * 65536);
  for (int a = 0; a < 65536; ++a) {
    float af = (float)(a * 2 * M_PI / 65536);  // int*int, then double (math.h:99)
    t[2 * a] = cosf(af);
    t[2 * a + 1] = sinf(af);
  }
  return t;
}

void make_rs_tables(uint8_t ex[512], uint8_t lg[256]) {
  // GF(256) mod x^8+x^4+x^3+x^2+1, alpha = x (rs.h:49-60, 89).  The reference
  // writes log[alpha^255 = 1] last, so log[1] == 255; log[0] is never written
  // there (uninitialised) and is pinned to 0 here.
  memset(ex, 0, 512);
  memset(lg, 0, 256);
  unsigned v = 1;
  for (unsigned i = 0; i < 256; ++i) {
    ex[i] = (uint8_t)v;
    ex[255 + i] = (uint8_t)v;
    lg[v] = (uint8_t)i;
    v <<= 1;
    if (v & 0x100) v ^= 0x11d;
  }
}

std::vector<uint8_t> make_derand_pattern() {
  // PRBS 1 + x^14 + x^15, initial state 100101010000000 (dvb.h:1116-1129):
  // 8 packets of 188 bytes, sync positions masked, first sync re-inverted.
  std::vector<uint8_t> p(1504);
  uint16_t st = 0251;  // octal, dvb.h:1119
  p[0] = 0xff;
  for (int i = 1; i < 1504; ++i) {
    uint8_t byte = 0;
    for (int k = 0; k < 8; ++k) {
      unsigned bit = ((st >> 13) ^ (st >> 14)) & 1;
      byte = (uint8_t)((byte << 1) | bit);
      st = (uint16_t)((st << 1) | bit);
    }
    p[i] = (i % 188) ? byte : 0;
  }
  return p;
}

namespace {

void dc_normalise(std::vector<float> &c, float gain) {  // filtergen.h:35-40
  float s = 0;
  for (float v : c) s = s + v;
  if (s) gain /= s;
  for (float &v : c) v = v * gain;
}

}  // namespace

std::vector<float> design_resampler(float Fs, float Fm, float rolloff, float rej,
                                    unsigned decim_opt, int *decim_out) {
  int decim;
  if (decim_opt) decim = (int)decim_opt;
  else {
    float target = Fm * 4;  // leandvb.cc:360-362
    decim = (int)(Fs / target);
    if (decim < 1) decim = 1;
  }
  float transition = (Fm / 2) * rolloff;         // leandvb.cc:363
  int order = (int)(rej * Fs / (22 * transition));
  order = ((order + 1) / 2) * 2;
  float Fcut = (Fm / 2) * (1 + rolloff / 2) / Fs;  // leandvb.cc:371
  int n = order + 1;
  std::vector<float> c(n);
  for (int i = 0; i < n; ++i) {  // filtergen.h:48-58: sinc, rectangular window
    float t = (float)(i - (n - 1) * 0.5);
    double arg = 2 * M_PI * Fcut * t;
    float sinc = (float)(2 * Fcut * (t ? sin(arg) / arg : 1));
    c[i] = sinc * 1.0f;
  }
  dc_normalise(c, 1);  // inside lowpass() (filtergen.h:59)
  dc_normalise(c, 1);  // again by the caller (leandvb.cc:376)
  *decim_out = decim;
  return c;
}

std::vector<float> shift_taps(const std::vector<float> &coeffs, float freq) {
  unsigned n = (unsigned)coeffs.size();
  std::vector<float> out(2 * n);
  for (unsigned i = 0; i < n; ++i) {
    unsigned k = i - n / 2;  // unsigned wrap for i < n/2, as in dsp.h:272
    float a = (float)(2 * M_PI * freq * k);
    out[2 * i] = coeffs[i] * cosf(a);
    out[2 * i + 1] = coeffs[i] * sinf(a);
  }
  return out;
}

// filtergen::root_raised_cosine (filtergen.h:68-92), float overloads of sin/cos/sqrt.
static std::vector<float> rrc_taps(int order, float fs, float rolloff) {
  float B = rolloff, pi = (float)M_PI;
  int n = (order + 1) | 1;  // filtergen.h:70
  std::vector<float> c(n);
  for (int i = 0; i < n; ++i) {
    int t = i - n / 2;
    float v;
    if (t == 0) v = sqrtf(fs) * (1 - B + 4 * B / pi);
    else {
      float tT = t * fs;
      float den = pi * tT * (1 - (4 * B * tT) * (4 * B * tT));
      if (!den)
        v = B * sqrtf(fs / 2) * ((1 + 2 / pi) * sinf(pi / (4 * B)) + (1 - 2 / pi) * cosf(pi / (4 * B)));
      else
        v = sqrtf(fs) * (sinf(pi * tT * (1 - B)) + 4 * B * tT * cosf(pi * tT * (1 + B))) / den;
    }
    c[i] = v;
  }
  dc_normalise(c, 1);
  return c;
}

std::vector<float> design_rrc(float Fs, float Fm, float rolloff, float rej,
                              int steps_opt, int *steps_out) {
  int steps = steps_opt;
  if (steps == 0) {  // leandvb.cc:442-445
    steps = (int)(64 * Fm / Fs);
    if (steps < 1) steps = 1;
  }
  float Frrc = Fs * steps;
  float transition = (Fm / 2) * rolloff;
  int order = (int)(rej * Frrc / (22 * transition));
  float fs = Fm / Frrc;  // symbol rate relative to the RRC sample rate
  std::vector<float> c = rrc_taps(order, fs, rolloff);
  *steps_out = steps;
  return c;
}

// leandvbtx.cc:131-138: interpolation RRC, then filtergen::normalize_power (filtergen.h:26-33).
std::vector<float> design_tx_rrc(int interp, float rolloff, float rrc_rej, float amp) {
  float Fm = 1.0 / interp;
  int order = interp * rrc_rej;
  std::vector<float> c = rrc_taps(order, Fm, rolloff);
  float gain = amp / 75.0f;   // cstln_amp (sdr.h:287)
  float s2 = 0;
  for (size_t i = 0; i < c.size(); ++i) s2 = s2 + c[i] * c[i];
  if (s2) gain /= sqrtf(s2);
  for (size_t i = 0; i < c.size(); ++i) c[i] = c[i] * gain;
  return c;
}

// fir_resampler::set_freq (dsp.h:352-361): no centring of the tap index here.
std::vector<float> shift_taps_resampler(const std::vector<float> &coeffs, float freq) {
  std::vector<float> out(2 * coeffs.size());
  for (size_t i = 0; i < coeffs.size(); ++i) {
    float a = (float)(2 * M_PI * freq * (int)i);
    out[2 * i] = coeffs[i] * cosf(a);
    out[2 * i + 1] = coeffs[i] * sinf(a);
  }
  return out;
}

std::vector<float> make_rotator_lut(float freq) {
  int ifreq = (int)(freq * 65536);  // sdr.h:1234
  std::vector<float> t(2 * 65536);
  for (int i = 0; i < 65536; ++i) {
    float a = (float)(2 * M_PI * i * ifreq / 65536);  // double, narrowed by cosf()
    t[i] = cosf(a);
    t[65536 + i] = sinf(a);
  }
  return t;
}

// ---------------------------------------------------------------- deconvolution

namespace {

struct Puncturer {
  uint32_t g[2], p[2];
  int period, weight;
  // IQ bit stream produced by feeding `s` MSB first (dvb.h:156-171).
  uint64_t encode(uint64_t s) const {
    uint64_t iq = 0;
    unsigned reg = 0;
    for (int b = bitlen(s) - 1; b >= 0; --b) {
      unsigned bit = (unsigned)(s >> b) & 1;
      reg = ((reg >> 1) | (bit << 6)) & 0xff;
      for (int j = 0; j < 2; ++j)
        if (p[j] & (1u << (b % period))) iq = (iq << 1) | par64(reg & g[j]);
    }
    return iq;
  }
};

// Smallest integer `d` with parity(d & resp[b]) == bit b of `want` for all 64
// columns, found by fixing bits from the LSB up and abandoning a branch as soon
// as a violated column has no tap left above the fixed bits (dvb.h:205-223).
void smallest_solution(const uint64_t resp[64], uint64_t fixed, int nfixed,
                       uint64_t want, uint64_t *best) {
  if (fixed > *best || nfixed > 64) return;
  bool all_ok = true;
  for (int b = 0; b < 64; ++b) {
    if (par64(fixed & resp[b]) != ((want >> b) & 1)) {
      if ((resp[b] >> (nfixed & 63)) == 0) return;
      all_ok = false;
    }
  }
  if (all_ok) { *best = fixed; return; }
  smallest_solution(resp, fixed, nfixed + 1, want, best);
  smallest_solution(resp, fixed | (1ull << (nfixed & 63)), nfixed + 1, want, best);
}

}  // namespace

// Alternate deconvolution polynomials for --fastlock: for every minimal polynomial d the
// reference pairs a second inverse d2 = d ^ (a parity check of the punctured code), chosen by
// its author and listed as constants (dvb.h:236-262).  They are data, reproduced here as
// {d, d2} pairs; make_deconv() verifies each one against the code (it must invert the encoder
// exactly like d does), so a wrong entry cannot go unnoticed.
static const uint64_t kAltPolys[][2] = {
    {0x3baull, 0x38ccaull},                                                          // 1/2
    {0xf29ull, 0x3c569329ull}, {0x3c552ull, 0x1dee1cull}, {0x7948ull, 0x1e2b49948ull}, {0x1deull, 0x1e2a90ull},   // 2/3
    {0xf247ull, 0xfd6383bull}, {0xfd9eeull, 0xfd91392ull}, {0xf248d8ull, 0xfd9eef18ull},                          // 3/4
    {0xf5727full, 0x3d5c909758full}, {0x3d5c90aaull, 0x0f5727f0229c90aaull}, {0x3daa371cull, 0x3d5f45630ecull},
    {0xf5727ff48ull, 0xf57d28260348ull}, {0xf57d28260ull, 0xf5727ff48128260ull},                                  // 5/6
    {0xfbeac76c454full, 0xfb11d6ba045a8full}, {0xfb11d6baull, 0xfbea3c7d930e16baull},
    {0xfb112d5038dcull, 0xfb112d5038271cull}, {0xfbea3c7d68ull, 0xfbeac7975462a8ull},
    {0xfb112d50ull, 0xfbea3c86793290ull}, {0xfb112dabd2e0ull, 0xfb112d50c3cd20ull},
    {0xfb11d640ull, 0xfbea3c8679c980ull},                                                                         // 7/8
};

bool make_deconv(int fec, DeconvPolys *out) {
  Puncturer pc;
  pc.g[0] = 0171; pc.g[1] = 0133;  // dvb.h:83-84
  switch (fec) {                   // dvb.h:486-511
    case 0: pc.p[0] = 0x1; pc.p[1] = 0x1; break;
    case 1: case 2: pc.p[0] = 0xa; pc.p[1] = 0xf; break;
    case 3: pc.p[0] = 0x5; pc.p[1] = 0x6; break;
    case 4: pc.p[0] = 0x15; pc.p[1] = 0x1a; break;
    case 5: pc.p[0] = 0x45; pc.p[1] = 0x7a; break;
    default: return false;
  }
  pc.period = bitlen(pc.p[0]) > bitlen(pc.p[1]) ? bitlen(pc.p[0]) : bitlen(pc.p[1]);
  pc.weight = __builtin_popcount(pc.p[0]) + __builtin_popcount(pc.p[1]);
  uint64_t resp[64];
  for (int b = 0; b < 64; ++b) resp[b] = pc.encode(1ull << b);
  out->punctperiod = pc.period;
  out->punctweight = pc.weight;
  for (int b = 0; b < pc.period; ++b) {
    uint64_t best = ~0ull;
    smallest_solution(resp, 0, 0, 1ull << b, &best);
    out->deconv[b] = best;
    // Known-answer self check (dvb.h:274-292): the polynomial must recover
    // bit b and nothing else from the response to every single input bit.
    for (int i = 0; i < 64; ++i)
      if (par64(resp[i] & best) != (unsigned)(b == i)) return false;
    uint64_t alt = best;
    for (const auto &pr : kAltPolys)
      if (pr[0] == best) alt = pr[1];
    if (alt == best) return false;   // dvb.h:263: "Alt polynomial not provided"
    for (int i = 0; i < 64; ++i)
      if (par64(resp[i] & alt) != (unsigned)(b == i)) return false;
    out->deconv2[b] = alt;
  }
  // Hypotheses {0, 90 degrees} x {direct, conjugate} (dvb.h:309-360); indexed by
  // the constellation symbol: bit1 of the symbol = (re<0), bit0 = (im<0) for
  // QPSK as laid out in sdr.h:333-336; the reference indexes
  // lut[(symbol&2)?1:0][symbol&1].
  for (int h = 0; h < 4; ++h)
    for (int a = 0; a <= 1; ++a)
      for (int b = 0; b <= 1; ++b) {
        int I = 0, Q = 0;
        switch (h) {
          case 0: I = a ? 0 : 1; Q = b ? 0 : 1; break;
          case 1: I = b ? 0 : 1; Q = !a ? 0 : 1; break;
          case 2: I = a ? 0 : 1; Q = b ? 1 : 0; break;
          case 3: I = b ? 1 : 0; Q = !a ? 0 : 1; break;
        }
        out->hyp_lut[h][a * 2 + b] = (uint8_t)((I << 1) | Q);
      }
  return true;
}

HsTables make_hs_tables() {
  HsTables t;
  t.polar.resize(65536); t.rect.resize(65536); t.sincos.resize(65536);
  for (int i = 0; i < 256; ++i)
    for (int q = 0; q < 256; ++q) {
      // (s_angle)(double) then assigned to a u_angle; (int)hypotf() assigned to an unsigned char
      const uint16_t a = (uint16_t)(int16_t)(int32_t)(atan2f(q - 128, i - 128) * 65536 / (2 * M_PI));
      const uint8_t r = (uint8_t)(int)hypotf(i - 128, q - 128);
      t.polar[i * 256 + q] = (uint32_t)a | ((uint32_t)r << 16);
    }
  for (unsigned long a = 0; a < 65536; ++a) {
    float f = 2 * M_PI * a / 65536;
    const uint8_t re = (uint8_t)(128 + 75 * cosf(f)), im = (uint8_t)(128 + 75 * sinf(f));   // cstln_amp (sdr.h:287)
    t.sincos[a] = (uint16_t)(re | (im << 8));
  }
  for (int a = 0; a < 256; ++a)
    for (int r = 0; r < 256; ++r) {
      const uint8_t re = (uint8_t)(int)(128 + r * cos(2 * M_PI * a / 256));
      const uint8_t im = (uint8_t)(int)(128 + r * sin(2 * M_PI * a / 256));
      t.rect[a * 256 + r] = (uint16_t)(re | (im << 8));
    }
  return t;
}

// ------------------------------------------------------------------- Viterbi

bool make_trellis(int fec, Trellis *t) {
  static const uint16_t G1 = 0171, G2 = 0133;
  uint16_t g[8];
  switch (fec) {  // dvb.h:519-565, 1180-1212
    case 0: t->bits_in = 1; t->bits_out = 2; g[0] = G1; g[1] = G2;
      t->path_nbits = 1; t->path_depth = 32; t->path32 = true; break;
    case 1: t->bits_in = 2; t->bits_out = 3; g[0] = G1; g[1] = G2; g[2] = G2 << 1;
      t->path_nbits = 3; t->path_depth = 21; break;
    case 2: t->bits_in = 4; t->bits_out = 6;
      g[0] = G1; g[1] = G2; g[2] = G2 << 1; g[3] = G1 << 2; g[4] = G2 << 2; g[5] = G2 << 3;
      t->path_nbits = 4; t->path_depth = 16; break;
    case 3: t->bits_in = 3; t->bits_out = 4; g[0] = G1; g[1] = G2; g[2] = G2 << 1; g[3] = G1 << 2;
      t->path_nbits = 3; t->path_depth = 21; break;
    case 4: t->bits_in = 5; t->bits_out = 6;
      g[0] = G1; g[1] = G2; g[2] = G2 << 1; g[3] = G1 << 2; g[4] = G2 << 3; g[5] = G1 << 4;
      t->path_nbits = 5; t->path_depth = 12; break;
    case 5: t->bits_in = 7; t->bits_out = 8;
      g[0] = G1; g[1] = G2; g[2] = G2 << 1; g[3] = G2 << 2; g[4] = G2 << 3; g[5] = G1 << 4;
      g[6] = G2 << 5; g[7] = G1 << 6;
      t->path_nbits = 7; t->path_depth = 9; break;
    default: return false;
  }
  t->nus = 1 << t->bits_in;
  t->ncs = 1 << t->bits_out;
  t->pred.assign((size_t)64 * t->ncs, 65);
  t->us.assign((size_t)64 * t->ncs, 0);
  for (int s = 0; s < 64; ++s)
    for (int u = 0; u < t->nus; ++u) {
      // The uncoded symbol enters bit-reversed above the 6 state bits
      // (viterbi.h:69-73), every polynomial taps the widened register.
      unsigned urev = 0;
      for (int b = 1; b < t->nus; b <<= 1) if (u & b) urev |= (unsigned)(t->nus / 2 / b);
      uint64_t reg = (uint64_t)s | ((uint64_t)urev << 6);
      unsigned label = 0;
      for (int k = 0; k < t->bits_out; ++k) label = (label << 1) | par64(reg & g[k]);
      unsigned next = (unsigned)(reg / t->nus);
      size_t at = (size_t)next * t->ncs + label;
      if (t->pred[at] != 65) return false;  // "Invalid convolutional code" (viterbi.h:83-86)
      t->pred[at] = (uint8_t)s;
      t->us[at] = (uint8_t)u;
    }
  return true;
}

VitSyncs make_vitsyncs(const Cstln &c, const Trellis &t) {
  VitSyncs v;
  int bps = 0;
  while ((1 << bps) < c.nsymbols) ++bps;
  v.bps = bps;
  int nconj = (c.nsymbols == 2) ? 1 : 2;  // dvb.h:1253-1257
  int nrot = (c.nsymbols == 2 || c.nsymbols == 4) ? c.nrotations / 2 : c.nrotations;
  v.nshifts = t.bits_out / bps;
  v.nsyncs = nconj * nrot * v.nshifts;
  v.shift.resize(v.nsyncs);
  v.map.assign(v.nsyncs, std::vector<uint8_t>(c.nsymbols, 0));
  for (int s = 0; s < v.nsyncs; ++s) {
    int rot = s % nrot, conj = (s / nrot) % nconj;
    v.shift[s] = s / nrot / nconj;
    float angle = (float)(2 * M_PI * rot / c.nrotations);  // dvb.h:1290
    float ca = cosf(angle), sa = sinf(angle);
    for (int i = 0; i < c.nsymbols; ++i) {  // dvb.h:1336-1351
      int8_t I = c.sym_re[i], Q = c.sym_im[i];
      if (conj) Q = (int8_t)-Q;
      int8_t RI = (int8_t)(I * ca - Q * sa);
      int8_t RQ = (int8_t)(I * sa + Q * ca);
      v.map[s][i] = (uint8_t)c.cells[(size_t)(uint8_t)RI * 256 + (uint8_t)RQ].symbol;
    }
  }
  return v;
}

}  // namespace ldvb
