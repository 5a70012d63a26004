/* oracle/dvbs_oracle.c -- TEST INFRASTRUCTURE ONLY (see dvbs_oracle.h).
 *
 * Scalar CPU restatement of the leandvb DVB-S receive path.  Every function
 * cites the reference lines whose behaviour it restates (paths relative to
 * /root/reference/src/).  Arithmetic notes that matter for bit parity:
 *  - the reference is built for baseline x86-64: SSE2 scalar float, no FMA;
 *    this file is compiled with -ffp-contract=off and no fast-math;
 *  - C++ overload resolution picks sinf/cosf/sqrtf/fabsf for float arguments
 *    in the reference (libstdc++ <math.h>); double literals promote to double;
 *  - float->integer conversions truncate toward zero (cvttss2si).
 */
#include "dvbs_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

typedef struct { float re, im; } cf;

static inline cf cmul(cf a, cf b) {           /* math.h:38-41 */
  cf r = { a.re*b.re - a.im*b.im, a.re*b.im + a.im*b.re };
  return r;
}

static inline unsigned parity64(uint64_t x) { /* math.h:75-88 */
  x ^= x >> 32; x ^= x >> 16; x ^= x >> 8; x ^= x >> 4;
  return (0x6996u >> (x & 15)) & 1;
}

static inline int hamming8(uint8_t x) {       /* math.h:57-60 */
  static const int lut[16] = { 0,1,1,2,1,2,2,3,1,2,2,3,2,3,3,4 };
  return lut[x&15] + lut[x>>4];
}

static int ilog2(uint64_t x) {                /* math.h:90-94, dvb.h:150-154 */
  int n = -1;
  for ( ; x; ++n, x >>= 1 ) ;
  return n;
}

/* ================================================================ tables */

/* sdr.h:492-495 */
static void polar(float r, int n, float i, int8_t *re, int8_t *im) {
  float a = (float)((double)(i * 2) * M_PI / n);
  *re = (int8_t)(r * cosf(a) * 75.0f);
  *im = (int8_t)(r * sinf(a) * 75.0f);
}

/* sdr.h:497-504: four points at the angles a*pi */
static void polar2(orc_cstln *c, int i, float r, float a0, float a1, float a2, float a3) {
  float a[] = { a0, a1, a2, a3 };
  for ( int j = 0; j < 4; ++j ) {
    float phi = a[j] * M_PI;
    c->sym_re[i+j] = (int8_t)(r*cosf(phi)*75.0f);
    c->sym_im[i+j] = (int8_t)(r*sinf(phi)*75.0f);
  }
}

/* sdr.h:505-528 */
static void make_qam(orc_cstln *c, int n) {
  c->nrotations = 4;
  c->nsymbols = n;
  int m = sqrtl(n);
  float scale;
  {
    int q = m / 2;
    float avgpower = 2*(q*0.25+(q-1)*q/2+(q-1)*q*(2*q-1)/6) / q;
    scale = 1.0 / sqrtf(avgpower);
  }
  int s = 0;
  for ( int x = 0; x < m; ++x )
    for ( int y = 0; y < m; ++y ) {
      float I = x - (float)(m-1)/2;
      float Q = y - (float)(m-1)/2;
      c->sym_re[s] = (int8_t)(I * scale * 75.0f);
      c->sym_im[s] = (int8_t)(Q * scale * 75.0f);
      ++s;
    }
}

#define P(s, r, n, i) polar(r, n, i, &c->sym_re[s], &c->sym_im[s])

int orc_cstln_build2(orc_cstln *c, int kind, int fec, int harden) {
  memset(c, 0, sizeof(*c));
  /* make_dvbs2_constellation, dvb.h:45-81 (DVB-S code rates only) */
  float gamma1 = 1, gamma2 = 1, gamma3 = 1;
  switch ( kind ) {
  case ORC_16APSK:
    switch ( fec ) {
    case ORC_FEC23: case ORC_FEC46: gamma1 = 3.15; break;
    case ORC_FEC34: gamma1 = 2.85; break;
    case ORC_FEC56: gamma1 = 2.70; break;
    default: return -1;
    }
    break;
  case ORC_32APSK:
    switch ( fec ) {
    case ORC_FEC34: gamma1 = 2.84; gamma2 = 5.27; break;
    case ORC_FEC56: gamma1 = 2.64; gamma2 = 4.64; break;
    default: return -1;
    }
    break;
  case ORC_64APSKE:
    gamma1 = 2.4; gamma2 = 4.3; gamma3 = 7;
    break;
  default: break;
  }
  switch ( kind ) {
  case ORC_BPSK:                               /* sdr.h:315-327 */
    c->nrotations = 2; c->nsymbols = 2;
    P(0, 1, 8, 1); P(1, 1, 8, 5);
    break;
  case ORC_QPSK:                               /* sdr.h:328-339 */
    c->nrotations = 4; c->nsymbols = 4;
    P(0, 1, 4, 0.5f); P(1, 1, 4, 3.5f); P(2, 1, 4, 1.5f); P(3, 1, 4, 2.5f);
    break;
  case ORC_8PSK: {                             /* sdr.h:340-354 */
    static const int idx[8] = { 1, 0, 4, 5, 2, 7, 3, 6 };
    c->nrotations = 8; c->nsymbols = 8;
    for ( int s = 0; s < 8; ++s ) P(s, 1, 8, (float)idx[s]);
    break;
  }
  case ORC_16APSK: {                           /* sdr.h:355-381 */
    float r1 = sqrtf(4 / (1+3*gamma1*gamma1));
    float r2 = gamma1 * r1;
    c->nrotations = 4; c->nsymbols = 16;
    P(0, r2, 12, 1.5f);  P(1, r2, 12, 10.5f); P(2, r2, 12, 4.5f);  P(3, r2, 12, 7.5f);
    P(4, r2, 12, 0.5f);  P(5, r2, 12, 11.5f); P(6, r2, 12, 5.5f);  P(7, r2, 12, 6.5f);
    P(8, r2, 12, 2.5f);  P(9, r2, 12, 9.5f);  P(10, r2, 12, 3.5f); P(11, r2, 12, 8.5f);
    P(12, r1, 4, 0.5f);  P(13, r1, 4, 3.5f);  P(14, r1, 4, 1.5f);  P(15, r1, 4, 2.5f);
    break;
  }
  case ORC_32APSK: {                           /* sdr.h:382-424 */
    float r1 = sqrtf(8 / (1+3*gamma1*gamma1+4*gamma2*gamma2));
    float r2 = gamma1 * r1;
    float r3 = gamma2 * r1;
    c->nrotations = 4; c->nsymbols = 32;
    P(0, r2, 12, 1.5f);  P(1, r2, 12, 2.5f);  P(2, r2, 12, 10.5f); P(3, r2, 12, 9.5f);
    P(4, r2, 12, 4.5f);  P(5, r2, 12, 3.5f);  P(6, r2, 12, 7.5f);  P(7, r2, 12, 8.5f);
    P(8, r3, 16, 1);     P(9, r3, 16, 3);     P(10, r3, 16, 14);   P(11, r3, 16, 12);
    P(12, r3, 16, 6);    P(13, r3, 16, 4);    P(14, r3, 16, 9);    P(15, r3, 16, 11);
    P(16, r2, 12, 0.5f); P(17, r1, 4, 0.5f);  P(18, r2, 12, 11.5f); P(19, r1, 4, 3.5f);
    P(20, r2, 12, 5.5f); P(21, r1, 4, 1.5f);  P(22, r2, 12, 6.5f); P(23, r1, 4, 2.5f);
    P(24, r3, 16, 0);    P(25, r3, 16, 2);    P(26, r3, 16, 15);   P(27, r3, 16, 13);
    P(28, r3, 16, 7);    P(29, r3, 16, 5);    P(30, r3, 16, 8);    P(31, r3, 16, 10);
    break;
  }
  case ORC_64APSKE: {                          /* sdr.h:425-452 */
    float r1 = sqrtf(64 / (4+12*gamma1*gamma1+20*gamma2*gamma2+28*gamma3*gamma3));
    float r2 = gamma1 * r1;
    float r3 = gamma2 * r1;
    float r4 = gamma3 * r1;
    c->nrotations = 4; c->nsymbols = 64;
    polar2(c,  0, r4,  1.0/ 4,  7.0/ 4,  3.0/ 4,  5.0/ 4);
    polar2(c,  4, r4, 13.0/28, 43.0/28, 15.0/28, 41.0/28);
    polar2(c,  8, r4,  1.0/28, 55.0/28, 27.0/28, 29.0/28);
    polar2(c, 12, r1,  1.0/ 4,  7.0/ 4,  3.0/ 4,  5.0/ 4);
    polar2(c, 16, r4,  9.0/28, 47.0/28, 19.0/28, 37.0/28);
    polar2(c, 20, r4, 11.0/28, 45.0/28, 17.0/28, 39.0/28);
    polar2(c, 24, r3,  1.0/20, 39.0/20, 19.0/20, 21.0/20);
    polar2(c, 28, r2,  1.0/12, 23.0/12, 11.0/12, 13.0/12);
    polar2(c, 32, r4,  5.0/28, 51.0/28, 23.0/28, 33.0/28);
    polar2(c, 36, r3,  9.0/20, 31.0/20, 11.0/20, 29.0/20);
    polar2(c, 40, r4,  3.0/28, 53.0/28, 25.0/28, 31.0/28);
    polar2(c, 44, r2,  5.0/12, 19.0/12,  7.0/12, 17.0/12);
    polar2(c, 48, r3,  1.0/ 4,  7.0/ 4,  3.0/ 4,  5.0/ 4);
    polar2(c, 52, r3,  7.0/20, 33.0/20, 13.0/20, 27.0/20);
    polar2(c, 56, r3,  3.0/20, 37.0/20, 17.0/20, 23.0/20);
    polar2(c, 60, r2,  1.0/ 4,  7.0/ 4,  3.0/ 4,  5.0/ 4);
    break;
  }
  case ORC_16QAM:  make_qam(c, 16);  break;
  case ORC_64QAM:  make_qam(c, 64);  break;
  case ORC_256QAM: make_qam(c, 256); break;
  default:
    fprintf(stderr, "orc_cstln_build: constellation not implemented\n");
    abort();
  }
  /* make_lut_from_symbols, sdr.h:526-561 */
  const int R = 256;
  for ( int I = -R/2; I < R/2; ++I )
    for ( int Q = -R/2; Q < R/2; ++Q ) {
      orc_cstln_cell *pr = &c->cell[I & (R-1)][Q & (R-1)];
      uint8_t nearest = 0;
      int32_t cost = R*R*2, cost2 = R*R*2;
      for ( int s = 0; s < c->nsymbols; ++s ) {
	int32_t d2 = (I - c->sym_re[s])*(I - c->sym_re[s]) +
	  (Q - c->sym_im[s])*(Q - c->sym_im[s]);
	if ( d2 < cost ) { cost2 = cost; cost = d2; nearest = s; }
	else if ( d2 < cost2 ) cost2 = d2;
      }
      if ( cost > 32767 ) cost = 32767;
      if ( cost2 > 32767 ) cost2 = 32767;
      pr->cost = (int16_t)(cost - cost2);
      pr->symbol = nearest;
      float ph_symbol = atan2f(c->sym_im[nearest], c->sym_re[nearest]);
      float ph_err = atan2f((float)Q, (float)I) - ph_symbol;
      /* (s32) is "signed long" = 64 bits on LP64; then stored modulo 2^16 */
      int64_t pe = (int64_t)((double)(ph_err * 65536) / (2*M_PI));
      pr->phase_error = (int16_t)(uint16_t)(uint64_t)pe;
      pr->pad = 0;
    }
  if ( harden )                                /* sdr.h:564-571 */
    for ( int i = 0; i < R; ++i )
      for ( int q = 0; q < R; ++q ) {
	if ( c->cell[i][q].cost < 0 ) c->cell[i][q].cost = -1;
	if ( c->cell[i][q].cost > 0 ) c->cell[i][q].cost = 1;
      }
  return 0;
}
#undef P

void orc_cstln_build(orc_cstln *c, int kind, int harden) {
  if ( orc_cstln_build2(c, kind, ORC_FEC34, harden) ) abort();
}

void orc_trig16_build(float *lut) {            /* math.h:97-103 */
  for ( int a = 0; a < 65536; ++a ) {
    float af = (float)(a * 2 * M_PI / 65536);
    lut[2*a] = cosf(af);
    lut[2*a+1] = sinf(af);
  }
}

/* rs.h:49-60.  lut_log[0] is never written by the reference; pinned to 0. */
static uint8_t gf_exp[512], gf_log[256], rs_G[17];
static int gf_ready = 0;

static inline uint8_t gf_mul(uint8_t x, uint8_t y) {  /* rs.h:64-67 */
  if ( !x || !y ) return 0;
  return gf_exp[gf_log[x] + gf_log[y]];
}
static inline uint8_t gf_div(uint8_t x, uint8_t y) {  /* rs.h:68-72 */
  if ( !x ) return 0;
  return gf_exp[gf_log[x] + 255 - gf_log[y]];
}
static inline uint8_t gf_inv(uint8_t x) {             /* rs.h:73-76 */
  return gf_exp[255 - gf_log[x]];
}

static void gf_init(void) {
  if ( gf_ready ) return;
  memset(gf_exp, 0, sizeof(gf_exp));
  memset(gf_log, 0, sizeof(gf_log));
  unsigned alpha_i = 1;
  for ( unsigned i = 0; i < 256; ++i ) {
    gf_exp[i] = alpha_i;
    gf_exp[255 + i] = alpha_i;
    gf_log[alpha_i] = i;             /* log[1] ends up 255 (i=255 overwrites) */
    alpha_i <<= 1;
    if ( alpha_i & 256 ) alpha_i ^= 0x11d;
  }
  /* Generator polynomial, rs.h:93-102 */
  for ( int i = 0; i <= 16; ++i ) rs_G[i] = (i == 16) ? 1 : 0;
  for ( int d = 0; d < 16; ++d )
    for ( int i = 0; i <= 16; ++i )
      rs_G[i] = ((i == 16) ? 0 : rs_G[i+1]) ^ gf_mul(gf_exp[d], rs_G[i]);
  gf_ready = 1;
}

void orc_rs_tables(uint8_t *exp511, uint8_t *log256, uint8_t *gen17) {
  gf_init();
  memcpy(exp511, gf_exp, 511);
  memcpy(log256, gf_log, 256);
  memcpy(gen17, rs_G, 17);
}

void orc_derand_pattern(uint8_t *pattern) {    /* dvb.h:1116-1129 */
  pattern[0] = 0xff;
  unsigned short st = 000251;
  for ( int i = 1; i < 188*8; ++i ) {
    uint8_t out = 0;
    for ( int n = 8; n--; ) {
      int bit = ((st >> 13) ^ (st >> 14)) & 1;
      out = (out << 1) | bit;
      st = (st << 1) | bit;
    }
    pattern[i] = (i % 188) ? out : 0;
  }
}

static void normalize_dcgain(int n, float *coeffs, float gain) { /* filtergen.h:35-40 */
  float s = 0;
  for ( int i = 0; i < n; ++i ) s = s + coeffs[i];
  if ( s ) gain /= s;
  for ( int i = 0; i < n; ++i ) coeffs[i] = coeffs[i] * gain;
}

int orc_lowpass(int order, float Fcut, float *coeffs) {  /* filtergen.h:45-62 */
  int ncoeffs = order + 1;
  for ( int i = 0; i < ncoeffs; ++i ) {
    float t = (float)(i - (ncoeffs-1)*0.5);
    float sinc = (float)(2*Fcut * (t ? sin(2*M_PI*Fcut*t)/(2*M_PI*Fcut*t) : 1));
    float window = 1;
    coeffs[i] = sinc * window;
  }
  normalize_dcgain(ncoeffs, coeffs, 1);
  return ncoeffs;
}

/* leandvb.cc:353-384 */
int orc_resample_design(float Fs, float Fm, float rolloff, float rej,
			unsigned decim_opt, float *coeffs, int max_coeffs,
			int *decim_out) {
  int decim;
  if ( decim_opt ) decim = decim_opt;
  else {
    float target_Fs = Fm * 4;
    decim = (int)(Fs / target_Fs);
    if ( decim < 1 ) decim = 1;
  }
  float transition = (Fm/2) * rolloff;
  int order = (int)(rej * Fs / (22*transition));
  order = ((order+1)/2) * 2;
  if ( order + 1 > max_coeffs ) return -1;
  float Fcut = (Fm/2) * (1 + rolloff/2) / Fs;
  int n = orc_lowpass(order, Fcut, coeffs);
  normalize_dcgain(n, coeffs, 1);
  *decim_out = decim;
  return n;
}

int orc_rrc(int order, float Fs, float rolloff, float *coeffs) { /* filtergen.h:68-92 */
  float B = rolloff, pi = (float)M_PI;
  int ncoeffs = (order+1) | 1;
  for ( int i = 0; i < ncoeffs; ++i ) {
    int t = i - ncoeffs/2;
    float c;
    if ( t == 0 )
      c = sqrtf(Fs) * (1 - B + 4*B/pi);
    else {
      float tT = t * Fs;
      float den = pi*tT*(1 - (4*B*tT)*(4*B*tT));
      if ( !den )
	c = B*sqrtf(Fs/2) * ( (1+2/pi)*sinf(pi/(4*B)) + (1-2/pi)*cosf(pi/(4*B)) );
      else
	c = sqrtf(Fs) * ( sinf(pi*tT*(1-B)) + 4*B*tT*cosf(pi*tT*(1+B)) ) / den;
    }
    coeffs[i] = c;
  }
  normalize_dcgain(ncoeffs, coeffs, 1);
  return ncoeffs;
}

/* ------------------------------------------- deconvolution polynomials */

static const uint32_t DVBS_G1 = 0171, DVBS_G2 = 0133;   /* dvb.h:83-84 */

typedef struct {
  uint32_t conv[2], punct[2];
  int punctperiod, punctweight;
  uint64_t response[64];
} dpoly_ctx;

static uint64_t dp_convolve(const dpoly_ctx *c, uint64_t s) {  /* dvb.h:156-171 */
  int sbits = ilog2(s) + 1;
  uint64_t iq = 0;
  unsigned char state = 0;
  for ( int b = sbits-1; b >= 0; --b ) {
    unsigned char bit = (s >> b) & 1;
    state = (state >> 1) | (bit << 6);
    for ( int j = 0; j < 2; ++j ) {
      unsigned char xy = parity64(state & c->conv[j]);
      if ( c->punct[j] & (1 << (b % c->punctperiod)) )
	iq = (iq << 1) | xy;
    }
  }
  return iq;
}

static void dp_solve_rec(const dpoly_ctx *c, uint64_t prefix, int nprefix,
			 uint64_t exp, uint64_t *best) {        /* dvb.h:205-223 */
  if ( prefix > *best ) return;
  if ( nprefix > 64 ) return;
  int solved = 1;
  for ( int b = 0; b < 64; ++b ) {
    if ( parity64(prefix & c->response[b]) != ((exp >> b) & 1) ) {
      /* shifting a 64-bit value by 64 is what the reference does at
	 nprefix==64 (x86: shift count masked); keep the masked behaviour */
      if ( (c->response[b] >> (nprefix & 63)) == 0 ) return;
      solved = 0;
    }
  }
  if ( solved ) { *best = prefix; return; }
  dp_solve_rec(c, prefix, nprefix+1, exp, best);
  dp_solve_rec(c, prefix | ((uint64_t)1 << (nprefix & 63)), nprefix+1, exp, best);
}

static void fec_punct(int fec, uint32_t *pX, uint32_t *pY) {  /* dvb.h:486-511 */
  switch ( fec ) {
  case ORC_FEC12: *pX = 0x1; *pY = 0x1; break;
  case ORC_FEC23:
  case ORC_FEC46: *pX = 0xa; *pY = 0xf; break;
  case ORC_FEC34: *pX = 0x5; *pY = 0x6; break;
  case ORC_FEC56: *pX = 0x15; *pY = 0x1a; break;
  case ORC_FEC78: *pX = 0x45; *pY = 0x7a; break;
  default: *pX = *pY = 1;
  }
}

static uint64_t alt_poly(uint64_t d) {         /* dvb.h:236-262 */
  static const uint64_t tab[][2] = {
    { 0x00000000000003baULL, 0x0000000000038ccaULL },
    { 0x0000000000000f29ULL, 0x000000003c569329ULL },
    { 0x000000000003c552ULL, 0x00000000001dee1cULL },
    { 0x0000000000007948ULL, 0x00000001e2b49948ULL },
    { 0x00000000000001deULL, 0x00000000001e2a90ULL },
    { 0x000000000000f247ULL, 0x000000000fd6383bULL },
    { 0x00000000000fd9eeULL, 0x000000000fd91392ULL },
    { 0x0000000000f248d8ULL, 0x00000000fd9eef18ULL },
    { 0x0000000000f5727fULL, 0x000003d5c909758fULL },
    { 0x000000003d5c90aaULL, 0x0f5727f0229c90aaULL },
    { 0x000000003daa371cULL, 0x000003d5f45630ecULL },
    { 0x0000000f5727ff48ULL, 0x0000f57d28260348ULL },
    { 0x0000000f57d28260ULL, 0xf5727ff48128260ULL },
    { 0x0000fbeac76c454fULL, 0x00fb11d6ba045a8fULL },
    { 0x00000000fb11d6baULL, 0xfbea3c7d930e16baULL },
    { 0x0000fb112d5038dcULL, 0x00fb112d5038271cULL },
    { 0x000000fbea3c7d68ULL, 0x00fbeac7975462a8ULL },
    { 0x00000000fb112d50ULL, 0x00fbea3c86793290ULL },
    { 0x0000fb112dabd2e0ULL, 0x00fb112d50c3cd20ULL },
    { 0x00000000fb11d640ULL, 0x00fbea3c8679c980ULL },
  };
  uint64_t d2 = d;
  for ( unsigned i = 0; i < sizeof(tab)/sizeof(tab[0]); ++i )
    if ( d == tab[i][0] ) d2 = tab[i][1];
  return d2;
}

int orc_deconv_polys(int fec, uint64_t *deconv, uint64_t *deconv2, int *punctweight) {
  dpoly_ctx c;
  c.conv[0] = DVBS_G1; c.conv[1] = DVBS_G2;
  fec_punct(fec, &c.punct[0], &c.punct[1]);
  c.punctperiod = 0; c.punctweight = 0;
  for ( int i = 0; i < 2; ++i ) {             /* dvb.h:139-143 */
    int nbits = ilog2(c.punct[i]) + 1;
    if ( nbits > c.punctperiod ) c.punctperiod = nbits;
    for ( uint32_t x = c.punct[i]; x; x >>= 1 ) c.punctweight += x & 1;
  }
  for ( int sbit = 0; sbit < 64; ++sbit )     /* dvb.h:227-230 */
    c.response[sbit] = dp_convolve(&c, (uint64_t)1 << sbit);
  for ( int b = 0; b < c.punctperiod; ++b ) {
    deconv[b] = ~(uint64_t)0;
    dp_solve_rec(&c, 0, 0, (uint64_t)1 << b, &deconv[b]);
    deconv2[b] = alt_poly(deconv[b]);
  }
  /* sanity check, dvb.h:274-292 */
  for ( int b = 0; b < c.punctperiod; ++b )
    for ( int i = 0; i < 64; ++i ) {
      uint64_t iq = dp_convolve(&c, (uint64_t)1 << i);
      unsigned expect = (b == i) ? 1 : 0;
      if ( parity64(iq & deconv[b]) != expect ||
	   parity64(iq & deconv2[b]) != expect ) {
	fprintf(stderr, "orc_deconv_polys: inverse convolution check failed\n");
	abort();
      }
    }
  *punctweight = c.punctweight;
  return c.punctperiod;
}

/* ============================================================ front end */

void orc_cconvert(const void *in, int fmt, float *out, size_t n) {  /* dsp.h:40-50 */
  size_t m = 2*n;
  switch ( fmt ) {
  case 0: { const uint8_t *p = in;             /* <u8,128,f32,0,1,1> */
      /* (pin->re-(Tin)Zin): u8 128 promotes to int, so int subtraction */
      for ( size_t i = 0; i < m; ++i ) out[i] = (float)(0 + ((int)p[i] - 128)*1/1);
      break; }
  case 1: { const int8_t *p = in;              /* <s8,0,...> */
      for ( size_t i = 0; i < m; ++i ) out[i] = (float)(0 + ((int)p[i] - 0)*1/1);
      break; }
  case 2: { const uint16_t *p = in;            /* <u16,32768,...> */
      for ( size_t i = 0; i < m; ++i ) out[i] = (float)(0 + ((int)p[i] - 32768)*1/1);
      break; }
  case 3: { const int16_t *p = in;             /* <s16,0,...> */
      for ( size_t i = 0; i < m; ++i ) out[i] = (float)(0 + ((int)p[i] - 0)*1/1);
      break; }
  default: abort();
  }
}

void orc_scale(const float *in, float scale, float *out, size_t n) { /* dsp.h:149-156 */
  for ( size_t i = 0; i < 2*n; ++i ) out[i] = in[i] * scale;
}

void orc_rotator_init(orc_rotator *r, float freq) {   /* sdr.h:1231-1241 */
  int ifreq = (int)(freq * 65536);
  for ( int i = 0; i < 65536; ++i ) {
    r->lut_cos[i] = cosf((float)(2*M_PI * i * ifreq / 65536));
    r->lut_sin[i] = sinf((float)(2*M_PI * i * ifreq / 65536));
  }
  r->index = 0;
}

void orc_rotator_run(orc_rotator *r, const float *in, float *out, size_t n) {
  for ( size_t k = 0; k < n; ++k, ++r->index ) {      /* sdr.h:1242-1254 */
    float c = r->lut_cos[r->index], s = r->lut_sin[r->index];
    float re = in[2*k], im = in[2*k+1];
    out[2*k]   = re*c - im*s;
    out[2*k+1] = re*s + im*c;
  }
}

void orc_fir_init(orc_fir *f, unsigned ncoeffs, const float *coeffs, unsigned decim) {
  f->ncoeffs = ncoeffs; f->coeffs = coeffs; f->decim = decim;
  f->shifted = malloc(sizeof(float)*2*ncoeffs);
  orc_fir_set_freq(f, 0);
}

void orc_fir_set_freq(orc_fir *f, float freq) {       /* dsp.h:270-280 */
  for ( int i = 0; i < (int)f->ncoeffs; ++i ) {
    /* (i-ncoeffs/2) is evaluated in UNSIGNED arithmetic in the reference */
    unsigned k = (unsigned)i - f->ncoeffs/2;
    float a = (float)(2*M_PI * freq * k);
    float c = cosf(a), s = sinf(a);
    f->shifted[2*i]   = f->coeffs[i] * c;
    f->shifted[2*i+1] = f->coeffs[i] * s;
  }
  f->current_freq = freq;
}

size_t orc_fir_run(orc_fir *f, const float *in, size_t n_in, float *out,
		   size_t *consumed) {                 /* dsp.h:233-262 */
  *consumed = 0;
  if ( n_in < f->ncoeffs ) return 0;
  size_t count = (n_in - f->ncoeffs) / f->decim;
  const cf *x = (const cf*)in;
  const cf *sc = (const cf*)f->shifted;
  cf *y = (cf*)out;
  for ( size_t k = 0; k < count; ++k ) {
    const cf *pi = x + f->ncoeffs + k*f->decim;
    cf acc = { 0, 0 };
    for ( unsigned i = 0; i < f->ncoeffs; ++i, --pi ) {
      cf p = cmul(sc[i], *pi);
      acc.re = acc.re + p.re;
      acc.im = acc.im + p.im;
    }
    y[k] = acc;
  }
  *consumed = count * f->decim;
  return count;
}

size_t orc_decimate(const float *in, size_t n_in, unsigned d, float *out,
		    size_t *consumed) {                /* generic.h:254-261 */
  size_t count = n_in / d;
  for ( size_t k = 0; k < count; ++k ) {
    out[2*k] = in[2*k*d]; out[2*k+1] = in[2*k*d+1];
  }
  *consumed = count * d;
  return count;
}

/* ------------------------------------------------------------ auto_notch */

static int *g_bitrev[17];
static float *g_omega_rev[17], *g_omega[17];

static void fft_tables(int n, int **bitrev, float **om, float **omrev) {
  int logn = 0;
  for ( int t = n; t > 1; t >>= 1 ) ++logn;
  if ( !g_bitrev[logn] ) {                     /* dsp.h:58-77 */
    int *br = malloc(sizeof(int)*n);
    for ( int i = 0; i < n; ++i ) {
      br[i] = 0;
      for ( int b = 0; b < logn; ++b ) br[i] = (br[i] << 1) | ((i >> b) & 1);
    }
    float *o = malloc(sizeof(float)*2*n), *orv = malloc(sizeof(float)*2*n);
    for ( int i = 0; i < n; ++i ) {
      float a = (float)(2.0*M_PI * i / n);
      orv[2*i]   =  (o[2*i]   = cosf(a));
      orv[2*i+1] = -(o[2*i+1] = sinf(a));
    }
    g_bitrev[logn] = br; g_omega[logn] = o; g_omega_rev[logn] = orv;
  }
  *bitrev = g_bitrev[logn]; *om = g_omega[logn]; *omrev = g_omega_rev[logn];
}

void orc_fft_inplace(int n, float *dataf, int reverse) {  /* dsp.h:78-110 */
  int *bitrev; float *om_f, *om_r;
  fft_tables(n, &bitrev, &om_f, &om_r);
  cf *data = (cf*)dataf;
  int logn = 0;
  for ( int t = n; t > 1; t >>= 1 ) ++logn;
  for ( int i = 0; i < n; ++i ) {
    int r = bitrev[i];
    if ( r < i ) { cf tmp = data[i]; data[i] = data[r]; data[r] = tmp; }
  }
  const cf *om = (const cf*)(reverse ? om_r : om_f);
  for ( int i = 0; i < logn; ++i ) {
    int hbs = 1 << i;
    int dom = 1 << (logn-1-i);
    for ( int j = 0; j < dom; ++j ) {
      int p = j*hbs*2, q = p + hbs;
      for ( int k = 0; k < hbs; ++k ) {
	cf w = om[k*dom];
	cf d = data[q+k];
	cf x = { w.re*d.re - w.im*d.im, w.re*d.im + w.im*d.re };
	data[q+k].re = data[p+k].re - x.re;
	data[q+k].im = data[p+k].im - x.im;
	data[p+k].re = data[p+k].re + x.re;
	data[p+k].im = data[p+k].im + x.im;
      }
    }
  }
  if ( reverse ) {
    float invn = (float)(1.0 / n);
    for ( int i = 0; i < n; ++i ) { data[i].re *= invn; data[i].im *= invn; }
  }
}

/* ------------------------------------------------------ cnr_fft / spectrum */

void orc_meas_init(orc_meas *m, int n, float bandwidth, float kavg, int decimation) {
  memset(m, 0, sizeof(*m));
  m->n = n; m->bandwidth = bandwidth; m->kavg = kavg; m->decimation = decimation;
}

static float meas_avgslots(const orc_meas *m, int i0, int i1) {   /* sdr.h:1333-1337 */
  float s = 0;
  for ( int i = i0; i <= i1; ++i ) s += m->avgpower[i & (m->n-1)];
  return s / (i1-i0+1);
}

size_t orc_meas_run(orc_meas *m, const float *in, size_t n_in, float center_freq, float *out,
		    size_t cap_points, size_t *consumed) {
  size_t points = 0, pos = 0;
  const int n = m->n;
  float *data = malloc(sizeof(float)*2*n), *power = malloc(sizeof(float)*n);
  while ( n_in - pos >= (size_t)n && points < cap_points ) {     /* sdr.h:1294-1302, 1362-1370 */
    m->phase += n;
    if ( m->phase >= m->decimation ) {
      m->phase -= m->decimation;
      memcpy(data, in + 2*pos, sizeof(float)*2*n);
      orc_fft_inplace(n, data, 1);
      for ( int i = 0; i < n; ++i )
	power[i] = data[2*i]*data[2*i] + data[2*i+1]*data[2*i+1];
      if ( !m->have_avg ) { memcpy(m->avgpower, power, sizeof(float)*n); m->have_avg = 1; }
      for ( int i = 0; i < n; ++i )
	m->avgpower[i] = m->avgpower[i]*(1-m->kavg) + power[i]*m->kavg;
      if ( m->bandwidth > 0 ) {                                  /* do_cnr, sdr.h:1306-1331 */
	int icf = (int)floor(center_freq*n+0.5);
	int bwslots = (int)((m->bandwidth/4) * n);
	if ( bwslots ) {
	  float c2plusn2 = meas_avgslots(m, icf-bwslots, icf+bwslots);
	  float n2 = ( meas_avgslots(m, icf-bwslots*4, icf-bwslots*3) +
		       meas_avgslots(m, icf+bwslots*3, icf+bwslots*4) ) / 2;
	  float c2 = c2plusn2 - n2;
	  float cnr = (c2>0 && n2>0) ? 10 * logf(c2/n2)/logf(10) : -50;
	  out[points++] = cnr;
	}
      } else {                                                   /* do_spectrum, sdr.h:1390-1396 */
	float *row = out + points*(size_t)n;
	for ( int i = 0; i < n/2; ++i ) {
	  row[i] = 10 * log10f(m->avgpower[n/2+i]);
	  row[n/2+i] = 10 * log10f(m->avgpower[i]);
	}
	++points;
      }
    }
    pos += n;
  }
  free(data); free(power);
  *consumed = pos;
  return points;
}

void orc_notch_init(orc_notch *a, int nslots) {       /* sdr.h:53-63 */
  memset(a, 0, sizeof(*a));
  a->nslots = nslots;
  a->decimation = 1024*4096;
  a->k = 0.002f;
  a->phase = 0;
  a->gain = 1;
  a->agc_rms_setpoint = 0;                     /* leandvb.cc:300-301 */
  for ( int s = 0; s < nslots; ++s ) {
    a->slots[s].i = -1;
    /* the reference leaves expj/estim uninitialised (fresh heap = zeros) */
    a->slots[s].expj = calloc(2*ORC_NOTCH_N, sizeof(float));
    a->slots[s].estim_re = a->slots[s].estim_im = 0;
  }
}

static void notch_detect(orc_notch *a, const float *pinf) {  /* sdr.h:76-118 */
  const int N = ORC_NOTCH_N;
  const cf *pin = (const cf*)pinf;
  static float data[2*ORC_NOTCH_N], amp[ORC_NOTCH_N];
  float m0 = 0, m2 = 0;
  for ( int i = 0; i < N; ++i ) {
    data[2*i] = pin[i].re; data[2*i+1] = pin[i].im;
    m2 += (float)pin[i].re*pin[i].re + (float)pin[i].im*pin[i].im;
    if ( fabsf(pin[i].re) > m0 ) m0 = fabsf(pin[i].re);
    if ( fabsf(pin[i].im) > m0 ) m0 = fabsf(pin[i].im);
  }
  if ( a->agc_rms_setpoint && m2 ) {
    float rms = sqrtf(m2/N);
    float new_gain = a->agc_rms_setpoint / rms;
    a->gain = (float)(a->gain*0.9 + new_gain*0.1);
  }
  orc_fft_inplace(N, data, 1);
  for ( int i = 0; i < N; ++i ) amp[i] = hypotf(data[2*i], data[2*i+1]);
  for ( int s = 0; s < a->nslots; ++s ) {
    int iamax = 0;
    for ( int i = 0; i < N; ++i ) if ( amp[i] > amp[iamax] ) iamax = i;
    if ( iamax != a->slots[s].i ) {
      a->slots[s].i = iamax;
      a->slots[s].estim_re = 0;
      a->slots[s].estim_im = 0;
      for ( int i = 0; i < N; ++i ) {
	float ang = (float)(2 * M_PI * a->slots[s].i * i / N);
	a->slots[s].expj[2*i]   = cosf(ang);
	a->slots[s].expj[2*i+1] = sinf(ang);
      }
    }
    amp[iamax] = 0;
    if ( iamax-1 >= 0 ) amp[iamax-1] = 0;
    if ( iamax+1 < N ) amp[iamax+1] = 0;
  }
}

static void notch_process(orc_notch *a, const float *pinf, float *poutf) { /* sdr.h:119-138 */
  const int N = ORC_NOTCH_N;
  const cf *pin = (const cf*)pinf;
  cf *pout = (cf*)poutf;
  float k = a->k;
  for ( int n = 0; n < N; ++n ) {
    cf out = pin[n];
    for ( int s = 0; s < a->nslots; ++s ) {
      float ejr = a->slots[s].expj[2*n], eji = a->slots[s].expj[2*n+1];
      float bbr = pin[n].re*ejr + pin[n].im*eji;
      float bbi = -pin[n].re*eji + pin[n].im*ejr;
      a->slots[s].estim_re = bbr*k + a->slots[s].estim_re*(1-k);
      a->slots[s].estim_im = bbi*k + a->slots[s].estim_im*(1-k);
      float subr = a->slots[s].estim_re*ejr - a->slots[s].estim_im*eji;
      float subi = a->slots[s].estim_re*eji + a->slots[s].estim_im*ejr;
      out.re -= subr;
      out.im -= subi;
    }
    pout[n].re = a->gain * out.re;
    pout[n].im = a->gain * out.im;
  }
}

size_t orc_notch_run(orc_notch *a, const float *in, size_t n_in, float *out) {
  size_t done = 0;                             /* sdr.h:64-75 */
  while ( n_in - done >= ORC_NOTCH_N ) {
    a->phase += ORC_NOTCH_N;
    if ( a->phase >= a->decimation ) {
      a->phase -= a->decimation;
      notch_detect(a, in + 2*done);
    }
    notch_process(a, in + 2*done, out + 2*done);
    done += ORC_NOTCH_N;
  }
  return done;
}

/* ============================================================= receiver */

static inline cf trig_expi(const float *trig, float a) {    /* math.h:104-110 */
  uint16_t idx = (uint16_t)(int16_t)(int32_t)a;
  cf r = { trig[2*idx], trig[2*idx+1] };
  return r;
}

static void rx_update_freq_limits(orc_rx *r) {        /* sdr.h:755-770 */
  int n = 4;
  if ( r->cstln ) {
    switch ( r->cstln->nsymbols ) {
    case 2: n = 2; break;
    case 4: n = 4; break;
    case 8: n = 8; break;
    case 16: n = 12; break;
    case 32: n = 16; break;
    default: n = 4; break;
    }
  }
  r->min_freqw = r->freqw - 65536/r->max_omega/n/2;
  r->max_freqw = r->freqw + 65536/r->max_omega/n/2;
}

void orc_rx_set_omega(orc_rx *r, float omega) {       /* sdr.h:738-743 */
  float tol = 10e-6;
  r->omega = omega;
  r->min_omega = omega * (1-tol);
  r->max_omega = omega * (1+tol);
  rx_update_freq_limits(r);
}

void orc_rx_set_freq(orc_rx *r, float freq) {         /* sdr.h:745-749 */
  r->freqw = freq * 65536;
  rx_update_freq_limits(r);
  r->freq_tap = r->freqw / 65536;
}

void orc_rx_init(orc_rx *r, const orc_cstln *c, const float *trig, int sampler) {
  memset(r, 0, sizeof(*r));                    /* sdr.h:709-736 */
  r->cstln = NULL;                             /* constructor runs with cstln NULL */
  r->trig = trig;
  r->sampler = sampler;
  r->meas_decimation = 1048576;
  r->pll_adjustment = 1.0f;
  r->allow_drift = 0;
  r->kest = 0.01f;
  r->est_insp = 75.0f*75.0f;
  r->agc_gain = 1;
  r->mu = 0; r->phase = 0; r->est_sp = 0; r->est_ep = 0; r->meas_count = 0;
  orc_rx_set_omega(r, 1);
  orc_rx_set_freq(r, 0);
  r->cstln = c;                                /* leandvb.cc:476 */
  r->samp_freqw = 0;   /* linear_sampler::freqw is set before first use */
}

void orc_rx_set_rrc(orc_rx *r, int ncoeffs, const float *coeffs, int subsampling) {
  r->rrc_ncoeffs = ncoeffs; r->rrc_coeffs = coeffs; r->rrc_sub = subsampling;
  r->rrc_shifted = calloc(2*ncoeffs, sizeof(float));
  r->rrc_update_phase = 0;                     /* sdr.h:639-643 */
}

int orc_rx_readahead(const orc_rx *r) {
  switch ( r->sampler ) {
  case ORC_SAMP_NEAREST: return 0;             /* sdr.h:594 */
  case ORC_SAMP_LINEAR: return 1;              /* sdr.h:607 */
  default: return r->rrc_ncoeffs - 1;          /* sdr.h:645 */
  }
}

static void rx_sampler_update_freq(orc_rx *r, float freqw) {
  if ( r->sampler == ORC_SAMP_LINEAR ) { r->samp_freqw = freqw; return; } /* sdr.h:620 */
  if ( r->sampler != ORC_SAMP_RRC ) return;
  r->rrc_update_phase -= 128;                  /* sdr.h:667-675 */
  if ( r->rrc_update_phase <= 0 ) {
    r->rrc_update_phase = r->rrc_ncoeffs*16;
    float f = freqw / r->rrc_sub;              /* sdr.h:678-682 */
    for ( int i = 0; i < r->rrc_ncoeffs; ++i ) {
      cf e = trig_expi(r->trig, -f*(i - r->rrc_ncoeffs/2));
      r->rrc_shifted[2*i]   = e.re * r->rrc_coeffs[i];
      r->rrc_shifted[2*i+1] = e.im * r->rrc_coeffs[i];
    }
  }
}

static inline cf rx_interp(orc_rx *r, const cf *pin, float mu, float phase) {
  switch ( r->sampler ) {
  case ORC_SAMP_NEAREST:                       /* sdr.h:595-597 */
    return cmul(pin[0], trig_expi(r->trig, -phase));
  case ORC_SAMP_LINEAR: {                      /* sdr.h:609-618 */
    cf s0 = cmul(pin[0], trig_expi(r->trig, -phase));
    cf s1 = cmul(pin[1], trig_expi(r->trig, -(phase + r->samp_freqw)));
    float a = 1 - mu;
    cf res = { s0.re*a + s1.re*mu, s0.im*a + s1.im*mu };
    return res;
  }
  default: {                                   /* sdr.h:647-665 */
    cf acc = { 0, 0 };
    const cf *sh = (const cf*)r->rrc_shifted;
    const cf *pc = sh + (int)((1-mu)*r->rrc_sub);
    const cf *pcend = sh + r->rrc_ncoeffs;
    for ( ; pc < pcend; pc += r->rrc_sub, ++pin ) {
      cf p = cmul(*pc, *pin);
      acc.re += p.re; acc.im += p.im;
    }
    return cmul(trig_expi(r->trig, -phase), acc);
  }
  }
}

size_t orc_rx_run(orc_rx *r, const float *inf, size_t n_in,
		  uint8_t *symbols_out, size_t *n_symbols,
		  float *sampled_out, size_t *n_sampled,
		  float *meas_out, size_t *n_meas) {   /* sdr.h:772-915 */
  const unsigned chunk_size = 128;
  const float cstln_amp = 75;
  float freq_alpha = 0.04;
  float freq_beta = 0.0012 / r->omega * r->pll_adjustment;
  float gain_mu = 0.02 / (cstln_amp*cstln_amp) * 2;
  const cf *in = (const cf*)inf;
  size_t done = 0, nsym = 0, nsamp = 0, nmeas = 0;
  int ra = orc_rx_readahead(r);

  while ( n_in - done >= chunk_size + ra ) {
    rx_sampler_update_freq(r, r->freqw);
    const cf *pin = in + done, *pend = pin + chunk_size;
    cf sg = { 0, 0 }, s = { 0, 0 };
    int have_point = 0;
    int8_t cp_re = 0, cp_im = 0;

    while ( pin < pend ) {
      if ( r->mu < 1 ) {
	sg = rx_interp(r, pin, r->mu, r->phase);
	s.re = sg.re * r->agc_gain;
	s.im = sg.im * r->agc_gain;
	/* cstln_lut::lookup(float,float), sdr.h:470-486 */
	float I = s.re, Q = s.im;
	while ( I < -128 || I > 127 || Q < -128 || Q > 127 ) { I *= 0.5; Q *= 0.5; }
	const orc_cstln_cell *cr =
	  &r->cstln->cell[(uint8_t)(int8_t)I][(uint8_t)(int8_t)Q];
	symbols_out[4*nsym+0] = (uint8_t)(cr->cost & 0xff);
	symbols_out[4*nsym+1] = (uint8_t)((cr->cost >> 8) & 0xff);
	symbols_out[4*nsym+2] = (uint8_t)cr->symbol;
	symbols_out[4*nsym+3] = 0;
	++nsym;
	/* PLL, sdr.h:814-816 */
	r->phase += cr->phase_error * freq_alpha;
	r->freqw += cr->phase_error * freq_beta;
	/* Modified Mueller and Muller, sdr.h:818-840 */
	r->hist[2] = r->hist[1];
	r->hist[1] = r->hist[0];
	r->hist[0].p_re = s.re;
	r->hist[0].p_im = s.im;
	cp_re = r->cstln->sym_re[cr->symbol];
	cp_im = r->cstln->sym_im[cr->symbol];
	have_point = 1;
	r->hist[0].c_re = cp_re;
	r->hist[0].c_im = cp_im;
	float muerr =
	  ( (r->hist[0].p_re - r->hist[2].p_re)*r->hist[1].c_re +
	    (r->hist[0].p_im - r->hist[2].p_im)*r->hist[1].c_im ) -
	  ( (r->hist[0].c_re - r->hist[2].c_re)*r->hist[1].p_re +
	    (r->hist[0].c_im - r->hist[2].c_im)*r->hist[1].p_im );
	float mucorr = muerr * gain_mu;
	const float max_mucorr = 0.1;
	if ( mucorr < -max_mucorr ) mucorr = -max_mucorr;
	if ( mucorr >  max_mucorr ) mucorr =  max_mucorr;
	r->mu += mucorr;
	r->mu += r->omega;
      }
      ++pin;
      --r->mu;
      r->phase += r->freqw;
    }
    done += chunk_size;

    r->phase = fmodf(r->phase, 65536);         /* sdr.h:855 */

    if ( have_point ) {
      if ( sampled_out ) { sampled_out[2*nsamp] = s.re; sampled_out[2*nsamp+1] = s.im; }
      ++nsamp;
      float insp = sg.re*sg.re + sg.im*sg.im;  /* sdr.h:863-869 */
      r->est_insp = insp*r->kest + r->est_insp*(1 - r->kest);
      if ( r->est_insp ) r->agc_gain = cstln_amp / sqrtf(r->est_insp);
      float evr = s.re - cp_re, evi = s.im - cp_im;   /* sdr.h:871-888 */
      float sig_power, ev_power;
      if ( r->cstln->nsymbols == 2 ) {
	float sig_real = (float)((cp_re + cp_im) * 0.707);
	float ev_real = (float)((evr + evi) * 0.707);
	sig_power = sig_real * sig_real;
	ev_power = ev_real * ev_real;
      } else {
	sig_power = (float)((int)cp_re*cp_re + (int)cp_im*cp_im);
	ev_power = evr*evr + evi*evi;
      }
      r->est_sp = sig_power*r->kest + r->est_sp*(1 - r->kest);
      r->est_ep = ev_power*r->kest + r->est_ep*(1 - r->kest);
    }

    if ( !r->allow_drift ) {                   /* sdr.h:895-898 */
      if ( r->freqw < r->min_freqw || r->freqw > r->max_freqw )
	r->freqw = (r->max_freqw + r->min_freqw) / 2;
    }

    r->freq_tap = r->freqw / 65536;            /* sdr.h:902, 917-919 */

    r->meas_count += chunk_size;               /* sdr.h:904-913 */
    while ( r->meas_count >= r->meas_decimation ) {
      r->meas_count -= r->meas_decimation;
      if ( meas_out ) {
	meas_out[3*nmeas+0] = r->freq_tap;
	meas_out[3*nmeas+1] = sqrtf(r->est_insp);
	meas_out[3*nmeas+2] = r->est_ep ? 10*logf(r->est_sp/r->est_ep)/logf(10) : 0;
      }
      ++nmeas;
    }
  }
  *n_symbols = nsym;
  if ( n_sampled ) *n_sampled = nsamp;
  if ( n_meas ) *n_meas = nmeas;
  return done;
}

/* ================================================ deconvolution and sync */

static void deconv_init_syncs(orc_deconv *d) {        /* dvb.h:309-360 */
  for ( int sync_id = 0; sync_id < 4; ++sync_id ) {
    for ( int re_pos = 0; re_pos <= 1; ++re_pos )
      for ( int im_pos = 0; im_pos <= 1; ++im_pos ) {
	int re_neg = !re_pos;
	int I = 0, Q = 0;
	switch ( sync_id ) {
	case 0: I = re_pos ? 0 : 1; Q = im_pos ? 0 : 1; break;
	case 1: I = im_pos ? 0 : 1; Q = re_neg ? 0 : 1; break;
	case 2: I = re_pos ? 0 : 1; Q = im_pos ? 1 : 0; break;
	case 3: I = im_pos ? 1 : 0; Q = re_neg ? 0 : 1; break;
	}
	d->syncs[sync_id].lut[re_pos][im_pos] = (I << 1) | Q;
      }
    d->syncs[sync_id].in = 0;  d->syncs[sync_id].n_in = 0;
    d->syncs[sync_id].out = 0; d->syncs[sync_id].n_out = 0;
    d->syncs[sync_id].in2 = 0; d->syncs[sync_id].n_in2 = 0; d->syncs[sync_id].n_out2 = 0;
  }
}

void orc_deconv_init(orc_deconv *d, int fec) {        /* dvb.h:124-148 */
  memset(d, 0, sizeof(*d));
  d->punctperiod = orc_deconv_polys(fec, d->deconv, d->deconv2, &d->punctweight);
  deconv_init_syncs(d);
  d->locked = 0;
  d->skip = 0;
}

void orc_deconv_next_sync(orc_deconv *d) {            /* dvb.h:185-193 */
  ++d->locked;
  if ( d->locked == 4 ) { d->locked = 0; d->skip = 1; }
}

static inline uint8_t deconv_readbyte(orc_deconv *d, orc_dsync *s,
				      const uint8_t **pp) {   /* dvb.h:369-389 */
  const int traceback = 64;
  const uint8_t *p = *pp;
  while ( s->n_out < 8 ) {
    uint64_t iq = s->in;
    while ( s->n_in < traceback ) {
      uint8_t sym = p[2];
      uint8_t iqbits = s->lut[(sym & 2) ? 1 : 0][sym & 1];
      p += 4;
      iq = (iq << 2) | iqbits;
      s->n_in += 2;
    }
    s->in = iq;
    for ( int b = d->punctperiod-1; b >= 0; --b ) {
      uint8_t bit = parity64(iq & d->deconv[b]);
      s->out = (s->out << 1) | bit;
    }
    s->n_out += d->punctperiod;
    s->n_in -= d->punctweight;
  }
  uint8_t res = (s->out >> (s->n_out-8)) & 255;
  s->n_out -= 8;
  *pp = p;
  return res;
}

/* deconvol_sync::readerrors (dvb.h:391-412): disagreements between the deconvolution
   polynomials and their alternates on the auxiliary register. */
static inline unsigned long deconv_readerrors(orc_deconv *d, orc_dsync *s,
					      const uint8_t **pp) {
  const int traceback = 64;
  const uint8_t *p = *pp;
  unsigned long res = 0;
  while ( s->n_out2 < 8 ) {
    uint64_t iq = s->in2;
    while ( s->n_in2 < traceback ) {
      uint8_t sym = p[2];
      uint8_t iqbits = s->lut[(sym & 2) ? 1 : 0][sym & 1];
      p += 4;
      iq = (iq << 2) | iqbits;
      s->n_in2 += 2;
    }
    s->in2 = iq;
    for ( int b = d->punctperiod-1; b >= 0; --b ) {
      uint8_t bit  = parity64(iq & d->deconv[b]);
      uint8_t bit2 = parity64(iq & d->deconv2[b]);
      if ( bit2 != bit ) ++res;
    }
    s->n_out2 += d->punctperiod;
    s->n_in2 -= d->punctweight;
  }
  s->n_out2 -= 8;
  *pp = p;
  return res;
}

void orc_deconv_set_fastlock(orc_deconv *d, int on) { d->fastlock = on; }

size_t orc_deconv_run(orc_deconv *d, const uint8_t *symbols4, size_t n_in,
		      uint8_t *out, size_t out_cap, size_t *consumed) {
  return orc_deconv_run2(d, symbols4, n_in, out, out_cap, consumed, 0);
}

void orc_deconv_set(orc_deconv *d, int locked, int skip) { d->locked = locked; d->skip = skip; }

size_t orc_deconv_run2(orc_deconv *d, const uint8_t *symbols4, size_t n_in,
		       uint8_t *out, size_t out_cap, size_t *consumed, int big_batch) {
  /* dvb.h:414-467 with fastlock == false */
  size_t skipped = 0;
  if ( d->skip ) {
    if ( n_in < (size_t)d->skip ) { *consumed = 0; return 0; }  /* would underflow */
    skipped = d->skip; d->skip = 0;
  }
  size_t readable = n_in - skipped;
  *consumed = skipped;
  if ( readable < 64 ) return 0;
  long maxrd = (long)((readable-64) / (d->punctweight/2) * d->punctperiod / 8);
  long maxwr = (long)out_cap;
  long n = (maxrd < maxwr) ? maxrd : maxwr;
  if ( !n ) return 0;
  if ( n < 32 && (!big_batch || d->fastlock) ) return 0;
  if ( d->fastlock ) {                         /* dvb.h:428-454: try all sync alignments */
    unsigned long errors_best = 1 << 30;
    int best = 0;
    for ( int k = 0; k < 4; ++k ) {
      const uint8_t *p = symbols4 + 4*skipped;
      unsigned long errors = 0;
      for ( long c = n; c--; ) errors += deconv_readerrors(d, &d->syncs[k], &p);
      if ( errors < errors_best ) { errors_best = errors; best = k; }
    }
    if ( best != d->locked ) d->locked = best;
    if ( errors_best > (unsigned long)(n*8/3) ) d->skip = 1;
  }
  const uint8_t *pin = symbols4 + 4*skipped, *pin0 = pin;
  uint8_t *pout = out;
  orc_dsync *s = &d->syncs[d->locked];
  while ( n-- ) *pout++ = deconv_readbyte(d, s, &pin);
  *consumed = skipped + (size_t)(pin - pin0)/4;
  return (size_t)(pout - out);
}

void orc_mpegsync_init(orc_mpegsync *m) {             /* dvb.h:719-741 */
  memset(m, 0, sizeof(*m));
  m->scan_syncs = 8; m->want_syncs = 4; m->lock_timeout = 4;
  m->polarity = 0; m->bitphase = 0; m->synchronized = 0;
  m->next_sync_count = 0; m->report_state = 1;
  m->phase8 = -1;
  m->fastlock = 0; m->resync_period = 1; m->resync_phase = 0;
}

void orc_mpegsync_set_fastlock(orc_mpegsync *m, int fastlock, int resync_period) {
  m->fastlock = fastlock; m->resync_period = resync_period;
}

/* dvb.h:798-840.  tmp must hold 204*8 bytes.  Returns bytes to skip (>0) on
   lock, 0 if no lock in this window. */
static int mpegsync_search(orc_mpegsync *m, const uint8_t *in, uint8_t *tmp) {
  const int P = 204;
  int chunk = P * m->scan_syncs;
  const uint8_t *pin = in, *pend = pin + chunk;
  uint8_t *pout = tmp;
  unsigned short w = *pin++;
  for ( ; pin <= pend; ++pin, ++pout ) {
    w = (w << 8) | *pin;
    *pout = w >> m->bitphase;
  }
  for ( int i = 0; i < P; ++i ) {
    int nsyncs_p = 0, nsyncs_n = 0;
    int phase8_p = -1, phase8_n = -1;
    const uint8_t *p = &tmp[i];
    for ( int j = 0; j < m->scan_syncs; ++j, p += P ) {
      uint8_t b = *p;
      if ( b == 0x47 ) { ++nsyncs_p; phase8_n = (8-j) & 7; }
      if ( b == 0xb8 ) { ++nsyncs_n; phase8_p = (8-j) & 7; }
    }
    int nsyncs;
    if ( nsyncs_p > nsyncs_n ) { m->polarity = 0;    nsyncs = nsyncs_p; m->phase8 = phase8_p; }
    else                       { m->polarity = 0xff; nsyncs = nsyncs_n; m->phase8 = phase8_n; }
    if ( nsyncs >= m->want_syncs && m->phase8 >= 0 ) {
      if ( !i ) { i = P; m->phase8 = (m->phase8+1) & 7; }
      m->synchronized = 1;
      m->lock_timeleft = m->lock_timeout;
      m->locktime = 0;
      return i;
    }
  }
  return 0;
}

size_t orc_mpegsync_run(orc_mpegsync *m, orc_deconv *deconv,
			const uint8_t *in, size_t n_in,
			uint8_t *out, size_t out_cap, size_t *consumed,
			int *lock_out, size_t *n_lock,
			uint64_t *locktime_out, size_t *n_locktime) {
  return orc_mpegsync_run2(m, deconv, in, n_in, out, out_cap, consumed,
			   lock_out, n_lock, locktime_out, n_locktime, 0, NULL);
}

/* per_wrap != 0 selects the large-batch schedule documented in DESIGN.md: the
   sweep counter advances at every wrap of the bit phase (what the reference
   does with its default buffers, where one run() call never sees two wraps) and
   the call returns as soon as deconv->next_sync() fired (*switched = 1) so that
   the caller can restart the deconvolver at exactly that byte. */
size_t orc_mpegsync_run2(orc_mpegsync *m, orc_deconv *deconv,
			 const uint8_t *in, size_t n_in,
			 uint8_t *out, size_t out_cap, size_t *consumed,
			 int *lock_out, size_t *n_lock,
			 uint64_t *locktime_out, size_t *n_locktime,
			 int per_wrap, int *switched) {
  const int P = 204;
  if ( switched ) *switched = 0;
  size_t rd = 0, wr = 0, nl = 0, nlt = 0;
  if ( m->report_state ) {                     /* dvb.h:743-747 */
    if ( lock_out ) lock_out[nl] = 0;
    ++nl;
    m->report_state = 0;
  }
  if ( m->synchronized ) {                     /* run_decoding, dvb.h:842-874 */
    while ( n_in - rd >= (size_t)P+1 && out_cap - wr >= (size_t)P ) {
      const uint8_t *pin = in + rd, *pend = pin + P;
      uint8_t *pout = out + wr;
      unsigned short w = *pin++;
      for ( ; pin <= pend; ++pin, ++pout ) {
	w = (w << 8) | *pin;
	*pout = (w >> m->bitphase) ^ m->polarity;
      }
      rd += P;
      uint8_t syncbyte = out[wr];
      wr += P;
      ++m->locktime;
      if ( locktime_out ) locktime_out[nlt] = m->locktime;
      ++nlt;
      uint8_t expected = m->phase8 ? 0x47 : 0xb8;
      if ( syncbyte == expected ) m->lock_timeleft = m->lock_timeout;
      m->phase8 = (m->phase8+1) & 7;
      --m->lock_timeleft;
      if ( !m->lock_timeleft ) {
	m->synchronized = 0;
	m->next_sync_count = 0;
	if ( lock_out ) lock_out[nl] = 0;
	++nl;
	break;
      }
    }
  } else if ( m->fastlock ) {                  /* run_searching_fast, dvb.h:781-796 */
    int chunk = P * m->scan_syncs;
    uint8_t tmp[204*9];
    while ( n_in - rd >= (size_t)chunk+1 && out_cap - wr >= (size_t)chunk ) {
      if ( m->resync_phase == 0 ) {
	for ( m->bitphase = 0; m->bitphase <= 7; ++m->bitphase ) {
	  int skip = mpegsync_search(m, in + rd, tmp);
	  if ( skip ) {
	    rd += skip;
	    if ( lock_out ) lock_out[nl] = 1;
	    ++nl;
	    goto done;
	  }
	}
      }
      rd += P;
      if ( ++m->resync_phase >= m->resync_period ) m->resync_phase = 0;
    }
  } else {                                     /* run_searching, dvb.h:755-779 */
    int next_sync = 0;
    int chunk = P * m->scan_syncs;
    uint8_t tmp[204*9];
    while ( n_in - rd >= (size_t)chunk+1 && out_cap - wr >= (size_t)chunk ) {
      int skip = mpegsync_search(m, in + rd, tmp);
      if ( skip ) {
	rd += skip;
	if ( lock_out ) lock_out[nl] = 1;
	++nl;
	goto done;                             /* search_sync() returned true */
      }
      rd += chunk;
      ++m->bitphase;
      if ( m->bitphase == 8 ) {
	m->bitphase = 0;
	next_sync = 1;
	if ( per_wrap ) {
	  next_sync = 0;
	  if ( ++m->next_sync_count >= 3 ) {
	    m->next_sync_count = 0;
	    if ( deconv ) orc_deconv_next_sync(deconv);
	    if ( switched ) *switched = 1;
	    goto done;
	  }
	}
      }
    }
    if ( next_sync ) {
      ++m->next_sync_count;
      if ( m->next_sync_count >= 3 ) {
	m->next_sync_count = 0;
	if ( deconv ) orc_deconv_next_sync(deconv);
      }
    }
  }
 done:
  *consumed = rd;
  if ( n_lock ) *n_lock = nl;
  if ( n_locktime ) *n_locktime = nlt;
  return wr;
}

size_t orc_deinterleave(const uint8_t *in, size_t n_in, uint8_t *out,
			size_t *consumed) {           /* dvb.h:933-945 */
  size_t rd = 0, np = 0;
  while ( n_in - rd >= 17*11*12 + 204 ) {
    const uint8_t *pin = in + rd + 17*11*12, *pend = pin + 204;
    uint8_t *pout = out + 204*np;
    for ( int delay = 17*11; pin < pend;
	  ++pin, ++pout, delay = (delay-17+17*12) % (17*12) )
      *pout = pin[-delay*12];
    rd += 204;
    ++np;
  }
  *consumed = rd;
  return np;
}

/* ------------------------------------------------------------------- RS */

static uint8_t rs_eval_poly_rev(const uint8_t *poly, int n, uint8_t x) { /* rs.h:125-130 */
  uint8_t acc = 0;
  for ( int i = 0; i < n; ++i ) acc = gf_mul(acc, x) ^ poly[i];
  return acc;
}

static uint8_t rs_eval_poly(const uint8_t *poly, int deg, uint8_t x) {   /* rs.h:133-138 */
  uint8_t acc = 0;
  for ( ; deg >= 0; --deg ) acc = gf_mul(acc, x) ^ poly[deg];
  return acc;
}

static int rs_syndromes(const uint8_t *poly, uint8_t *synd) {            /* rs.h:116-123 */
  int corrupted = 0;
  for ( int i = 0; i < 16; ++i ) {
    synd[i] = rs_eval_poly_rev(poly, 204, gf_exp[i]);
    if ( synd[i] ) corrupted = 1;
  }
  return corrupted;
}

void orc_rs_encode(uint8_t *msg) {                   /* rs.h:142-170 */
  gf_init();
  uint8_t p[204];
  memcpy(p, msg, 188);
  memset(p+188, 0, 16);
  for ( int d = 0; d < 188; ++d ) {
    if ( !p[d] ) continue;
    uint8_t k = gf_div(p[d], rs_G[0]);
    for ( int i = 0; i <= 16; ++i ) p[d+i] ^= gf_mul(k, rs_G[i]);
  }
  memcpy(msg+188, p+188, 16);
}

static int rs_correct(uint8_t synd[16], uint8_t *pout, uint8_t *pin,
		      int *bits_corrected) {          /* rs.h:176-268 */
  /* One spare element: the reference reads C[16] out of bounds when L==16
     (needs S0..S14==0, S15!=0); here that element is a defined 0. */
  uint8_t C[17] = { 1 }, B[17] = { 1 };
  int L = 0, m = 1;
  uint8_t b = 1;
  for ( int n = 0; n < 16; ++n ) {
    uint8_t d = synd[n];
    for ( int i = 1; i <= L; ++i ) d ^= gf_mul(C[i], synd[n-i]);
    if ( !d ) {
      ++m;
    } else if ( 2*L <= n ) {
      uint8_t T[16];
      memcpy(T, C, 16);
      for ( int i = 0; i < 16-m; ++i )
	C[m+i] ^= gf_mul(d, gf_mul(gf_inv(b), B[i]));
      L = n + 1 - L;
      memcpy(B, T, 16);
      b = d;
      m = 1;
    } else {
      for ( int i = 0; i < 16-m; ++i )
	C[m+i] ^= gf_mul(d, gf_mul(gf_inv(b), B[i]));
      ++m;
    }
  }
  uint8_t omega[16];
  memset(omega, 0, sizeof(omega));
  for ( int i = 0; i < 16; ++i )
    for ( int j = 0; j < 16; ++j )
      if ( i+j < 16 ) omega[i+j] ^= gf_mul(synd[i], C[j]);
  uint8_t Cprime[15];
  for ( int i = 0; i < 15; ++i ) Cprime[i] = (i & 1) ? 0 : C[i+1];
  int roots_found = 0;
  for ( int i = 0; i < 255; ++i ) {
    uint8_t r = gf_exp[i];
    uint8_t v = rs_eval_poly(C, L, r);
    if ( !v ) {
      uint8_t xk = gf_inv(r);
      int loc = (255-i) % 255;
      if ( loc < 204 ) {
	uint8_t num = gf_mul(xk, rs_eval_poly(omega, L < 16 ? L : 15, r));
	uint8_t den = rs_eval_poly(Cprime, 14, r);
	uint8_t e = gf_div(num, den);
	if ( bits_corrected ) *bits_corrected += hamming8(e);
	if ( loc >= 16 ) pout[203-loc] ^= e;
	if ( pin ) pin[203-loc] ^= e;
      }
      if ( ++roots_found == L ) break;
    }
  }
  if ( pin ) return rs_syndromes(pin, synd);
  return 0;
}

int orc_rs_decode_packet(uint8_t *pin, uint8_t *pout, int *bits_corrected) {
  gf_init();                                   /* dvb.h:1004-1047 */
  memcpy(pout, pin, 188);
  uint8_t synd[16];
  int corrupted = rs_syndromes(pin, synd);
  if ( corrupted ) corrupted = rs_correct(synd, pout, pin, bits_corrected);
  if ( corrupted ) pout[0] ^= 0x55;
  return corrupted;
}

size_t orc_derandomize(orc_derand *d, const uint8_t *in, size_t npackets,
		       uint8_t *out) {                /* dvb.h:1131-1158 */
  static uint8_t pattern[188*8];
  static int ready = 0;
  if ( !ready ) { orc_derand_pattern(pattern); ready = 1; }
  size_t nout = 0;
  for ( size_t k = 0; k < npackets; ++k ) {
    const uint8_t *pin = in + 188*k;
    uint8_t *pout = out + 188*nout;
    if ( pin[0] == 0xb8 || pin[0] == (0xb8 ^ 0x55) ) d->pos = 0;
    for ( int i = 0; i < 188; ++i ) pout[i] = pin[i] ^ pattern[d->pos + i];
    d->pos += 188;
    if ( d->pos == 188*8 ) d->pos = 0;
    if ( pout[0] == 0x47 ) ++nout;
    /* else: TEI bit set on a packet that is never committed (dvb.h:1151) */
  }
  return nout;
}

/* ============================================================== Viterbi */

typedef struct { uint8_t pred, us; } vbranch;

typedef struct {
  int32_t cost[2][64];
  uint64_t path[2][64];
  int bank;
} vdec;

struct orc_viterbi {
  const orc_cstln *cstln;
  int bits_in, bits_out, bps;
  int nus, ncs;
  int path_nbits, path_depth, path32;
  vbranch *trellis;          /* [64][ncs] */
  int nsyncs, nshifts;
  struct { int shift; uint8_t map[256]; vdec dec; } *syncs;
  int current_sync, resync_phase, resync_period;
};

static const uint16_t polys12[] = { 0171, 0133 };                   /* dvb.h:519-548 */
static const uint16_t polys23[] = { 0171, 0133, 0133<<1 };
static const uint16_t polys46[] = { 0171, 0133, 0133<<1, 0171<<2, 0133<<2, 0133<<3 };
static const uint16_t polys34[] = { 0171, 0133, 0133<<1, 0171<<2 };
static const uint16_t polys56[] = { 0171, 0133, 0133<<1, 0171<<2, 0133<<3, 0171<<4 };
static const uint16_t polys78[] = { 0171, 0133, 0133<<1, 0133<<2, 0133<<3, 0171<<4,
				    0133<<5, 0171<<6 };

orc_viterbi *orc_viterbi_new(const orc_cstln *c, int fec) {
  orc_viterbi *v = calloc(1, sizeof(*v));
  const uint16_t *G;
  v->cstln = c;
  switch ( fec ) {            /* dvb.h:550-565, 1180-1212 */
  case ORC_FEC12: v->bits_in=1; v->bits_out=2; G=polys12; v->path_nbits=1; v->path_depth=32; v->path32=1; break;
  case ORC_FEC23: v->bits_in=2; v->bits_out=3; G=polys23; v->path_nbits=3; v->path_depth=21; break;
  case ORC_FEC46: v->bits_in=4; v->bits_out=6; G=polys46; v->path_nbits=4; v->path_depth=16; break;
  case ORC_FEC34: v->bits_in=3; v->bits_out=4; G=polys34; v->path_nbits=3; v->path_depth=21; break;
  case ORC_FEC56: v->bits_in=5; v->bits_out=6; G=polys56; v->path_nbits=5; v->path_depth=12; break;
  case ORC_FEC78: v->bits_in=7; v->bits_out=8; G=polys78; v->path_nbits=7; v->path_depth=9; break;
  default: free(v); return NULL;
  }
  v->nus = 1 << v->bits_in;
  v->ncs = 1 << v->bits_out;
  v->bps = ilog2(c->nsymbols);
  if ( v->bits_out % v->bps ) { free(v); return NULL; }
  /* trellis::init_convolutional, viterbi.h:61-92 */
  v->trellis = malloc(sizeof(vbranch)*64*v->ncs);
  for ( int i = 0; i < 64*v->ncs; ++i ) v->trellis[i].pred = 65;
  int nG = ilog2(v->ncs);
  for ( int s = 0; s < 64; ++s )
    for ( int us = 0; us < v->nus; ++us ) {
      uint64_t shiftreg = s;
      int us_rev = 0;
      for ( int b = 1; b < v->nus; b *= 2 ) if ( us & b ) us_rev |= (v->nus/2/b);
      shiftreg |= (uint64_t)us_rev * 64;
      uint32_t cs = 0;
      for ( int g = 0; g < nG; ++g ) cs = (cs << 1) | parity64(shiftreg & G[g]);
      shiftreg /= v->nus;
      vbranch *b = &v->trellis[shiftreg*v->ncs + cs];
      if ( b->pred != 65 ) { fprintf(stderr, "Invalid convolutional code\n"); abort(); }
      b->pred = s; b->us = us;
    }
  /* viterbi_sync constructor, dvb.h:1236-1297 */
  int nconj = (c->nsymbols == 2) ? 1 : 2;
  int nrot = (c->nsymbols == 2 || c->nsymbols == 4) ? c->nrotations/2 : c->nrotations;
  v->nshifts = v->bits_out / v->bps;
  v->nsyncs = nconj * nrot * v->nshifts;
  v->syncs = calloc(v->nsyncs, sizeof(*v->syncs));
  for ( int s = 0; s < v->nsyncs; ++s ) {
    int rot = s % nrot;
    int conj = (s/nrot) % nconj;
    int shift = s / nrot / nconj;
    v->syncs[s].shift = shift;
    /* init_map, dvb.h:1336-1351 */
    float angle = (float)(2*M_PI*rot/c->nrotations);
    float ca = cosf(angle), sa = sinf(angle);
    for ( int i = 0; i < c->nsymbols; ++i ) {
      int8_t I = c->sym_re[i], Q = c->sym_im[i];
      if ( conj ) Q = -Q;
      int8_t RI = (int8_t)(I*ca - Q*sa);
      int8_t RQ = (int8_t)(I*sa + Q*ca);
      v->syncs[s].map[i] = (uint8_t)c->cell[(uint8_t)RI][(uint8_t)RQ].symbol;
    }
    /* viterbi_dec constructor, viterbi.h:133-145: bank 0 costs = 0, paths = 0 */
  }
  v->current_sync = 0;
  v->resync_phase = 0;
  v->resync_period = 32;
  return v;
}

void orc_viterbi_free(orc_viterbi *v) {
  if ( !v ) return;
  free(v->trellis); free(v->syncs); free(v);
}
void orc_viterbi_set_resync_period(orc_viterbi *v, int p) { v->resync_period = p; }
int orc_viterbi_nsyncs(const orc_viterbi *v) { return v->nsyncs; }
int orc_viterbi_current_sync(const orc_viterbi *v) { return v->current_sync; }

/* viterbi_dec::update(1,&cs,&cost,quality), viterbi.h:202-260 */
static uint8_t vdec_update(orc_viterbi *v, vdec *d, uint8_t cs, int32_t cost,
			   int32_t *quality) {
  const int32_t max_tpm = INT32_MAX;
  int32_t best_tpm = max_tpm, best2_tpm = max_tpm;
  int best_state = 0;
  int cur = d->bank, nxt = cur ^ 1;
  for ( int s = 0; s < 64; ++s ) {
    int32_t best_m = max_tpm;
    const vbranch *best_b = NULL;
    const vbranch *row = &v->trellis[s*v->ncs];
    {
      const vbranch *b = &row[cs];
      if ( b->pred != 65 ) {
	int32_t m = d->cost[cur][b->pred] + cost;
	if ( m <= best_m ) { best_m = m; best_b = b; }
      }
    }
    if ( 1 != v->ncs ) {
      for ( int c = 0; c < v->ncs; ++c ) {
	const vbranch *b = &row[c];
	if ( b->pred == 65 ) continue;
	int32_t m = d->cost[cur][b->pred];
	if ( m <= best_m ) { best_m = m; best_b = b; }
      }
    }
    uint64_t p = d->path[cur][best_b->pred];
    if ( v->path32 ) p = (uint32_t)((uint32_t)p << v->path_nbits) | best_b->us;
    else p = (p << v->path_nbits) | best_b->us;
    d->path[nxt][s] = p;
    d->cost[nxt][s] = best_m;
    if ( best_m < best_tpm ) { best_state = s; best2_tpm = best_tpm; best_tpm = best_m; }
    else if ( best_m < best2_tpm ) best2_tpm = best_m;
  }
  d->bank = nxt;
  for ( int s = 0; s < 64; ++s ) d->cost[nxt][s] -= best_tpm;
  if ( quality ) *quality = best2_tpm - best_tpm;
  uint64_t p = d->path[nxt][best_state];
  return (uint8_t)((p >> ((v->path_depth-1)*v->path_nbits)) & ((1u << v->path_nbits)-1));
}

static uint8_t vit_update_sync(orc_viterbi *v, int s, const uint8_t *pin4,
			       int32_t *discr) {      /* dvb.h:1353-1364 */
  pin4 += 4*v->syncs[s].shift;
  uint8_t cs = 0;
  int32_t cost = 0;
  for ( int i = 0; i < v->nshifts; ++i, pin4 += 4 ) {
    cs = (uint8_t)((cs << v->bps) | v->syncs[s].map[pin4[2]]);
    cost += (int16_t)(pin4[0] | (pin4[1] << 8));
  }
  return vdec_update(v, &v->syncs[s].dec, cs, cost, discr);
}

size_t orc_viterbi_run(orc_viterbi *v, const uint8_t *symbols4, size_t n_in,
		       uint8_t *out, size_t out_cap, size_t *consumed) {
  const int chunk_size = 128;                  /* dvb.h:1366-1414 */
  int discr_delay = 64 / v->bits_in;
  size_t rd = 0, wr = 0;
  int32_t *totaldiscr = malloc(sizeof(int32_t)*v->nsyncs);
  while ( n_in - rd >= (size_t)(v->nshifts*chunk_size + (v->nshifts-1)) &&
	  (out_cap - wr)*8 >= (size_t)(v->bits_in*chunk_size) ) {
    for ( int s = 0; s < v->nsyncs; ++s ) totaldiscr[s] = 0;
    uint64_t outstream = 0;
    int nout = 0;
    const uint8_t *pin = symbols4 + 4*rd;
    for ( int blocknum = 0; blocknum < chunk_size; ++blocknum, pin += 4*v->nshifts ) {
      int32_t discr;
      uint8_t result = vit_update_sync(v, v->current_sync, pin, &discr);
      outstream = (outstream << v->bits_in) | result;
      nout += v->bits_in;
      if ( blocknum >= discr_delay ) totaldiscr[v->current_sync] += discr;
      if ( !v->resync_phase ) {
	for ( int s = 0; s < v->nsyncs; ++s ) {
	  if ( s == v->current_sync ) continue;
	  int32_t dsc;
	  (void)vit_update_sync(v, s, pin, &dsc);
	  if ( blocknum >= discr_delay ) totaldiscr[s] += dsc;
	}
      }
      while ( nout >= 8 ) {
	out[wr++] = (uint8_t)(outstream >> (nout-8));
	nout -= 8;
      }
    }
    rd += (size_t)chunk_size*v->nshifts;
    if ( !v->resync_phase ) {
      int best = v->current_sync;
      for ( int s = 0; s < v->nsyncs; ++s )
	if ( totaldiscr[s] > totaldiscr[best] ) best = s;
      v->current_sync = best;
    }
    if ( ++v->resync_phase >= v->resync_period ) v->resync_phase = 0;
  }
  free(totaldiscr);
  *consumed = rd;
  return wr;
}

/* ---- model-checking hooks for the time-segment schedule of the CUDA Viterbi stage (k_viterbi.cu):
   the same block update as above (pinned to the reference), driven one chunk at a time by the test. */

/* One chunk (128 FEC blocks) at symbols4 for the decoders in run_mask; returns the bytes of decoder
   `out_sync` (or nothing when out_sync < 0); totals[s] = sum of quality over blocks >= discr_delay. */
size_t orc_viterbi_chunk(orc_viterbi *v, const uint8_t *symbols4, uint32_t run_mask, int out_sync,
			 uint8_t *out, int32_t *totals) {
  const int chunk_size = 128;
  int discr_delay = 64 / v->bits_in;
  size_t wr = 0;
  for ( int s = 0; s < v->nsyncs; ++s ) {
    totals[s] = 0;
    if ( !((run_mask >> s) & 1) ) continue;
    uint64_t outstream = 0;
    int nout = 0;
    const uint8_t *pin = symbols4;
    for ( int blocknum = 0; blocknum < chunk_size; ++blocknum, pin += 4*v->nshifts ) {
      int32_t discr;
      uint8_t result = vit_update_sync(v, s, pin, &discr);
      if ( blocknum >= discr_delay ) totals[s] += discr;
      if ( s == out_sync ) {
	outstream = (outstream << v->bits_in) | result;
	nout += v->bits_in;
	while ( nout >= 8 ) { out[wr++] = (uint8_t)(outstream >> (nout-8)); nout -= 8; }
      }
    }
  }
  return wr;
}

/* Normalised metrics and path registers of one decoder, as the CUDA stage stores them (bank re-based). */
void orc_viterbi_get_dec(const orc_viterbi *v, int s, int32_t *cost64, uint64_t *path64) {
  const vdec *d = &v->syncs[s].dec;
  memcpy(cost64, d->cost[d->bank], sizeof(int32_t)*64);
  memcpy(path64, d->path[d->bank], sizeof(uint64_t)*64);
}
void orc_viterbi_set_dec(orc_viterbi *v, int s, const int32_t *cost64, const uint64_t *path64) {
  vdec *d = &v->syncs[s].dec;
  d->bank = 0;
  memcpy(d->cost[0], cost64, sizeof(int32_t)*64);
  memcpy(d->path[0], path64, sizeof(uint64_t)*64);
}
void orc_viterbi_set_ctl(orc_viterbi *v, int current_sync, int resync_phase) {
  v->current_sync = current_sync; v->resync_phase = resync_phase;
}
int orc_viterbi_resync_phase(const orc_viterbi *v) { return v->resync_phase; }
int orc_viterbi_nshifts(const orc_viterbi *v) { return v->nshifts; }
int orc_viterbi_bits_in(const orc_viterbi *v) { return v->bits_in; }

/* ===================================================== binding helpers */
/* Flat accessors so that the Python test harness (oracle/oracle.py) can
   allocate and inspect stage state without mirroring struct layouts. */

size_t orc_sizeof(int what) {
  switch ( what ) {
  case 0: return sizeof(orc_cstln);
  case 1: return sizeof(orc_rotator);
  case 2: return sizeof(orc_fir);
  case 3: return sizeof(orc_notch);
  case 4: return sizeof(orc_rx);
  case 5: return sizeof(orc_deconv);
  case 6: return sizeof(orc_mpegsync);
  case 7: return sizeof(orc_derand);
  case 8: return sizeof(orc_meas);
  case 9: return sizeof(orc_hsrx);
  case 10: return sizeof(orc_hsdeconv);
  default: return 0;
  }
}

/* 22 words: mu phase freqw est_insp agc_gain est_sp est_ep hist[3]{p_re,p_im,c_re,c_im}
   samp_freqw freq_tap meas_count(lo) */
void orc_rx_get_state(const orc_rx *r, uint32_t *w) {
  float f[21] = { r->mu, r->phase, r->freqw, r->est_insp, r->agc_gain, r->est_sp, r->est_ep,
		  r->hist[0].p_re, r->hist[0].p_im, r->hist[0].c_re, r->hist[0].c_im,
		  r->hist[1].p_re, r->hist[1].p_im, r->hist[1].c_re, r->hist[1].c_im,
		  r->hist[2].p_re, r->hist[2].p_im, r->hist[2].c_re, r->hist[2].c_im,
		  r->samp_freqw, r->freq_tap };
  memcpy(w, f, sizeof(f));
  w[21] = (uint32_t)r->meas_count;
}

void orc_rx_set_state(orc_rx *r, const uint32_t *w) {
  float f[21];
  memcpy(f, w, sizeof(f));
  r->mu = f[0]; r->phase = f[1]; r->freqw = f[2]; r->est_insp = f[3]; r->agc_gain = f[4];
  r->est_sp = f[5]; r->est_ep = f[6];
  for ( int k = 0; k < 3; ++k ) {
    r->hist[k].p_re = f[7+4*k]; r->hist[k].p_im = f[8+4*k];
    r->hist[k].c_re = f[9+4*k]; r->hist[k].c_im = f[10+4*k];
  }
  r->samp_freqw = f[19]; r->freq_tap = f[20];
  r->meas_count = w[21];
}

void orc_rx_config(orc_rx *r, float pll_adjustment, int allow_drift,
		   unsigned long meas_decimation) {
  r->pll_adjustment = pll_adjustment;
  r->allow_drift = allow_drift;
  r->meas_decimation = meas_decimation;
}

void orc_rx_get_limits(const orc_rx *r, float *f4) {
  f4[0] = r->min_freqw; f4[1] = r->max_freqw; f4[2] = r->omega; f4[3] = r->freq_tap;
}

/* notch state: phase, gain, then per slot {i, estim_re, estim_im} */
void orc_notch_get_state(const orc_notch *a, int32_t *phase, float *gain,
			 int32_t *slot_i, float *estim /* [nslots][2] */) {
  *phase = a->phase; *gain = a->gain;
  for ( int s = 0; s < a->nslots; ++s ) {
    slot_i[s] = a->slots[s].i;
    estim[2*s] = a->slots[s].estim_re; estim[2*s+1] = a->slots[s].estim_im;
  }
}

const float *orc_fir_shifted(const orc_fir *f) { return f->shifted; }
float orc_fir_current_freq(const orc_fir *f) { return f->current_freq; }

void orc_deconv_get(const orc_deconv *d, int *locked, int *skip,
		    int *punctperiod, int *punctweight) {
  *locked = d->locked; *skip = d->skip;
  *punctperiod = d->punctperiod; *punctweight = d->punctweight;
}

/* bitphase polarity synchronized phase8 next_sync_count lock_timeleft locktime */
void orc_mpegsync_get(const orc_mpegsync *m, int64_t *v7) {
  v7[0] = m->bitphase; v7[1] = m->polarity; v7[2] = m->synchronized; v7[3] = m->phase8;
  v7[4] = m->next_sync_count; v7[5] = (int64_t)m->lock_timeleft; v7[6] = (int64_t)m->locktime;
}
