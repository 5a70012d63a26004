/* oracle/dvbs_tx_oracle.c -- TEST INFRASTRUCTURE ONLY (see dvbs_oracle.h).
 *
 * CPU restatement (plain C, scalar) of the leandvbtx DVB-S transmit chain of
 * pabr/leansdr, stage by stage on whole arrays: every stage sees the complete
 * output of the previous one, which is the fixpoint the reference scheduler
 * reaches (apps/leandvbtx.cc:79-197).  Citations: /root/reference/src/<file>:<lines>.
 *
 * Parity pin: tests/test_oracle_tx_cpu.py compares the final cf32 stream bit
 * for bit with the output of the UNMODIFIED reference binary oracle/_ref/leandvbtx
 * (QPSK 1/2 and 7/8, 8PSK 2/3, with and without --agc), and the committed golden
 * hashes under tests/golden/tx_kat.json.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "dvbs_oracle.h"

/* ---- randomizer (leansdr/dvb.h:1063-1102): the pattern is the derandomizer's
 * (dvb.h:1116-1129 builds the same 1504 bytes), position restarts every 8 packets. */
void orc_tx_randomize(const uint8_t *ts, size_t npk, uint8_t *out) {
  uint8_t pattern[1504];
  orc_derand_pattern(pattern);
  size_t pos = 0;
  for ( size_t p = 0; p < npk; ++p ) {
    for ( int i = 0; i < 188; ++i, ++pos ) out[p*188+i] = ts[p*188+i] ^ pattern[pos];
    if ( pos == 1504 ) pos = 0;
  }
}

/* ---- rs_encoder (leansdr/dvb.h:957-980, rs.h:141-167) */
void orc_tx_rs_encode(const uint8_t *ts188, size_t npk, uint8_t *rs204) {
  for ( size_t p = 0; p < npk; ++p ) {
    memcpy(rs204 + p*204, ts188 + p*188, 188);
    orc_rs_encode(rs204 + p*204);
  }
}

/* ---- interleaver (leansdr/dvb.h:900-921): needs 12 packets readable, emits one
 * 204-byte row per packet consumed; the last 11 packets stay in the pipe. */
size_t orc_tx_interleave(const uint8_t *rs204, size_t npk, uint8_t *out) {
  size_t rows = npk >= 12 ? npk - 11 : 0;
  for ( size_t r = 0; r < rows; ++r ) {
    int delay = 0;
    for ( int i = 0; i < 204; ++i, delay = (delay+1) % 12 )
      out[r*204+i] = rs204[(r + 11 - delay)*204 + i];
  }
  return rows * 204;
}

/* ---- dvb_convol / convol_multipoly<uint16_t,16> (leansdr/dvb.h:519-604,
 * convolutional.h:225-270).  Consumes multiples of bits_in bytes. */
static const uint16_t G1 = 0171, G2 = 0133;          /* dvb.h:84-85 */
static int tx_polys(int fec, uint16_t *p, int *bits_in) {   /* dvb.h:520-565 */
  switch ( fec ) {
  case ORC_FEC12: p[0]=G1; p[1]=G2; *bits_in = 1; return 2;
  case ORC_FEC23: p[0]=G1; p[1]=G2; p[2]=G2<<1; *bits_in = 2; return 3;
  case ORC_FEC46: p[0]=G1; p[1]=G2; p[2]=G2<<1; p[3]=G1<<2; p[4]=G2<<2; p[5]=G2<<3; *bits_in = 4; return 6;
  case ORC_FEC34: p[0]=G1; p[1]=G2; p[2]=G2<<1; p[3]=G1<<2; *bits_in = 3; return 4;
  case ORC_FEC56: p[0]=G1; p[1]=G2; p[2]=G2<<1; p[3]=G1<<2; p[4]=G2<<3; p[5]=G1<<4; *bits_in = 5; return 6;
  case ORC_FEC78: p[0]=G1; p[1]=G2; p[2]=G2<<1; p[3]=G2<<2; p[4]=G2<<3; p[5]=G1<<4; p[6]=G2<<5; p[7]=G1<<6;
    *bits_in = 7; return 8;
  }
  return 0;
}
static inline int parity16(uint16_t x) { return __builtin_parity(x); }   /* math.h:62-73 */

size_t orc_tx_convol(int fec, int bps, const uint8_t *in, size_t n, uint8_t *out, size_t *consumed) {
  uint16_t polys[8]; int bits_in;
  const int bits_out = tx_polys(fec, polys, &bits_in);
  if ( !bits_out || bits_out % bps ) { *consumed = 0; return 0; }
  n = (n / bits_in) * bits_in;                         /* dvb.h:591-593 */
  uint16_t hist = 0, sersymb = 0; int nhist = 0, nsersymb = 0;
  const uint8_t symbmask = (uint8_t)((1 << bps) - 1);
  uint8_t *pout = out;
  for ( size_t k = 0; k < n; ++k ) {
    uint8_t b = in[k];
    for ( int bit = 8; bit--; ) {
      hist = (uint16_t)((hist >> 1) | ((uint16_t)((b >> bit) & 1) << 15));
      if ( ++nhist == bits_in ) {
	for ( int p = 0; p < bits_out; ++p )
	  sersymb = (uint16_t)((sersymb << 1) | parity16((uint16_t)(hist & polys[p])));
	nhist = 0;
	nsersymb += bits_out;
	while ( nsersymb >= bps ) {
	  *pout++ = (uint8_t)((sersymb >> (nsersymb - bps)) & symbmask);
	  nsersymb -= bps;
	}
      }
    }
  }
  *consumed = n;
  return (size_t)(pout - out);
}

/* ---- cstln_transmitter<f32,0> (leansdr/sdr.h:1196-1221) */
void orc_tx_map(const orc_cstln *c, const uint8_t *sym, size_t n, float *out_cf32) {
  for ( size_t i = 0; i < n; ++i ) {
    out_cf32[2*i]   = 0 + c->sym_re[sym[i]];
    out_cf32[2*i+1] = 0 + c->sym_im[sym[i]];
  }
}

/* ---- RRC interpolation taps (apps/leandvbtx.cc:131-138, filtergen.h:26-33,68-92) */
float orc_tx_amp(const char *power_db) {               /* leandvbtx.cc:289 */
  return expf(logf(10) * atof(power_db) / 20);
}
int orc_tx_taps(int interp, float rolloff, float rrc_rej, float amp, float *coeffs) {
  float Fm = 1.0 / interp;
  int order = interp * rrc_rej;
  int n = orc_rrc(order, Fm, rolloff, coeffs);
  float gain = amp / 75.0f;                            /* cstln_amp = 75 (sdr.h:287) */
  float s2 = 0;
  for ( int i = 0; i < n; ++i ) s2 = s2 + coeffs[i]*coeffs[i];
  if ( s2 ) gain /= sqrtf(s2);
  for ( int i = 0; i < n; ++i ) coeffs[i] = coeffs[i] * gain;
  return n;
}

/* ---- fir_resampler<cf32,float> (leansdr/dsp.h:290-364), decim = 1, freq 0.
 * Fixpoint of run(): nothing while fewer than ncoeffs items are readable, then
 * count = (readable*interp - ncoeffs)/interp input steps. */
size_t orc_tx_resample(const float *xin, size_t n_in, const float *coeffs, int ncoeffs, int interp,
		       float *yout, size_t *consumed) {
  *consumed = 0;
  if ( n_in < (size_t)ncoeffs || n_in*interp < (size_t)ncoeffs ) return 0;
  float *sc = malloc(sizeof(float) * 2 * ncoeffs);
  for ( int i = 0; i < ncoeffs; ++i ) {                /* set_freq(0), dsp.h:352-360 */
    float a = 2*M_PI * 0.0f * i;
    float c = cosf(a), s = sinf(a);
    sc[2*i] = coeffs[i] * c;
    sc[2*i+1] = coeffs[i] * s;
  }
  size_t count = (n_in*interp - ncoeffs) / interp;
  int latency = (ncoeffs + interp) / interp;
  const float *pin = xin + 2*latency;
  float *pout = yout;
  for ( size_t n = 0; n < count; ++n, pin += 2 ) {
    for ( int i = 0; i < interp; ++i, pout += 2 ) {
      const float *pi = pin;
      float xr = 0, xi = 0;
      for ( int pc = i; pc < ncoeffs; pc += interp, pi -= 2 ) {
	float cr = sc[2*pc], ci = sc[2*pc+1];
	float pr = cr*pi[0] - ci*pi[1];                /* math.h:38-41 */
	float pj = cr*pi[1] + ci*pi[0];
	xr = xr + pr; xi = xi + pj;
      }
      pout[0] = xr; pout[1] = xi;
    }
  }
  free(sc);
  *consumed = count;
  return count * interp;
}

/* ---- simple_agc<f32> (leansdr/sdr.h:238-274), chunks of 128 samples. */
size_t orc_tx_agc(const float *xin, size_t n, float out_rms, float bw, float *yout) {
  float estimated = 0;
  size_t done = 0;
  while ( n - done >= 128 ) {
    const float *pin = xin + 2*done;
    float amp2 = 0;
    for ( int i = 0; i < 128; ++i ) amp2 += pin[2*i]*pin[2*i] + pin[2*i+1]*pin[2*i+1];
    amp2 /= 128;
    if ( !estimated ) estimated = amp2;
    estimated = estimated*(1-bw) + amp2*bw;
    float gain = estimated ? out_rms / sqrtf(estimated) : 0;
    for ( int i = 0; i < 128; ++i ) {
      yout[2*(done+i)]   = pin[2*i]   * gain;
      yout[2*(done+i)+1] = pin[2*i+1] * gain;
    }
    done += 128;
  }
  return done;
}

/* ---- the whole chain (apps/leandvbtx.cc:79-197): TS packets -> cf32.  Returns samples. */
size_t orc_tx_chain(const uint8_t *ts, size_t npk, int cstln_kind, int fec, int interp, int decim,
		    float rolloff, float rrc_rej, const char *power_db, int agc,
		    float *out, size_t cap_samples,
		    uint8_t *tap_mpegbytes, size_t *n_mpegbytes, uint8_t *tap_symbols, size_t *n_symbols) {
  orc_cstln *c = malloc(sizeof *c);
  if ( orc_cstln_build2(c, cstln_kind, fec, 0) ) { free(c); return 0; }   /* leandvbtx.cc:112 */
  int bps = 0; while ( (1 << bps) < c->nsymbols ) ++bps;
  if ( fec == ORC_FEC23 && (c->nsymbols == 4 || c->nsymbols == 64) ) fec = ORC_FEC46;   /* leandvbtx.cc:115-119 */
  uint8_t *r = malloc(npk*188 + 1), *rs = malloc(npk*204 + 1), *mb = malloc(npk*204 + 1);
  orc_tx_randomize(ts, npk, r);
  orc_tx_rs_encode(r, npk, rs);
  size_t nmb = orc_tx_interleave(rs, npk, mb);
  uint8_t *sym = malloc(nmb*8*2 + 16);
  size_t used = 0;
  size_t nsym = orc_tx_convol(fec, bps, mb, nmb, sym, &used);
  if ( tap_mpegbytes ) memcpy(tap_mpegbytes, mb, nmb);
  if ( n_mpegbytes ) *n_mpegbytes = nmb;
  if ( tap_symbols ) memcpy(tap_symbols, sym, nsym);
  if ( n_symbols ) *n_symbols = nsym;
  float *iq = malloc(sizeof(float)*2*(nsym + 1));
  orc_tx_map(c, sym, nsym, iq);
  float amp = orc_tx_amp(power_db);
  float coeffs[4096];
  int order = interp * rrc_rej;
  size_t n_out = 0;
  if ( order + 2 <= 4096 ) {
    int nc = orc_tx_taps(interp, rolloff, rrc_rej, amp, coeffs);
    float *up = malloc(sizeof(float)*2*(nsym*interp + 1));
    size_t cons = 0;
    size_t nup = orc_tx_resample(iq, nsym, coeffs, nc, interp, up, &cons);
    size_t nd = nup / decim;                           /* decimator, generic.h:247-267 */
    float *dec = malloc(sizeof(float)*2*(nd + 1));
    for ( size_t i = 0; i < nd; ++i ) { dec[2*i] = up[2*i*decim]; dec[2*i+1] = up[2*i*decim+1]; }
    if ( agc ) {
      float out_rms = amp / sqrtf((float)interp/decim);     /* leandvbtx.cc:163-165 */
      float bw = 0.001 * decim / interp;
      float *y = malloc(sizeof(float)*2*(nd + 1));
      n_out = orc_tx_agc(dec, nd, out_rms, bw, y);
      if ( n_out > cap_samples ) n_out = cap_samples;
      memcpy(out, y, sizeof(float)*2*n_out);
      free(y);
    } else {
      n_out = nd > cap_samples ? cap_samples : nd;
      memcpy(out, dec, sizeof(float)*2*n_out);
    }
    free(up); free(dec);
  }
  free(iq); free(sym); free(r); free(rs); free(mb); free(c);
  return n_out;
}
