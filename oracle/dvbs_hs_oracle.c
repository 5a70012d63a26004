/* oracle/dvbs_hs_oracle.c -- TEST INFRASTRUCTURE ONLY (see dvbs_oracle.h).
 *
 * CPU restatement of the two runnables that are specific to leandvb's --hs path
 * (apps/leandvb.cc:727-969): fast_qpsk_receiver<u8> (leansdr/sdr.h:946-1189) and
 * dvb_deconvol_sync_hard = dvb_deconvol_sync<u8> over deconvol_poly2<u8,uint32_t,
 * uint64_t,0x3ba,0x38f70> (leansdr/dvb.h:612-707, convolutional.h:75-192).  The rest
 * of that path (mpeg_sync with fastlock, deinterleaver, rs_decoder, derandomizer) is
 * shared with the default path (dvbs_oracle.c).
 *
 * Parity pin: tests/test_oracle_cpu.py compares hard symbols, bytes and TS with the
 * UNMODIFIED reference (`leandvb --hs` and oracle/_ref/ref_hs, a harness that taps the
 * reference runnables).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "dvbs_oracle.h"

/* ------------------------------------------------------- fast_qpsk_receiver */

void orc_hsrx_init(orc_hsrx *r) {                      /* sdr.h:957-975, 1144-1164 */
  memset(r, 0, sizeof *r);
  r->meas_decimation = 1048576;
  r->pll_adjustment = 1.0f;
  r->allow_drift = 0;
  for ( int i = 0; i < 256; ++i )
    for ( int q = 0; q < 256; ++q ) {
      r->polar_a[i][q] = (uint16_t)(int16_t)(int32_t)(atan2f(q-128, i-128) * 65536 / (2*M_PI));
      r->polar_r[i][q] = (uint8_t)(int)hypotf(i-128, q-128);
    }
  for ( unsigned long a = 0; a < 65536; ++a ) {
    float f = 2*M_PI * a / 65536;
    r->sincos[a][0] = (uint8_t)(128 + 75*cosf(f));     /* cstln_amp = 75 (sdr.h:287) */
    r->sincos[a][1] = (uint8_t)(128 + 75*sinf(f));
  }
  for ( int a = 0; a < 256; ++a )
    for ( int k = 0; k < 256; ++k ) {
      r->rect[a][k][0] = (uint8_t)(int)(128 + k*cos(2*M_PI*a/256));
      r->rect[a][k][1] = (uint8_t)(int)(128 + k*sin(2*M_PI*a/256));
    }
  orc_hsrx_set_omega(r, 1);
  orc_hsrx_set_freq(r, 0);
}

static void hsrx_update_freq_limits(orc_hsrx *r) {     /* sdr.h:989-994 */
  r->min_freqw = r->freqw - 65536/r->max_omega/8;
  r->max_freqw = r->freqw + 65536/r->max_omega/8;
}

void orc_hsrx_set_omega(orc_hsrx *r, float omega) {    /* sdr.h:977-982 */
  float tol = 10e-6;
  r->omega = omega;
  r->min_omega = omega * (1-tol);
  r->max_omega = omega * (1+tol);
  hsrx_update_freq_limits(r);
}

void orc_hsrx_set_freq(orc_hsrx *r, float freq) {      /* sdr.h:984-987 */
  r->freqw = freq * 65536;
  hsrx_update_freq_limits(r);
}

void orc_hsrx_config(orc_hsrx *r, int allow_drift, unsigned long meas_decimation) {
  r->allow_drift = allow_drift;
  r->meas_decimation = meas_decimation;
}

/* sdr.h:999-1140.  in: interleaved u8 I,Q.  Returns samples consumed. */
size_t orc_hsrx_run(orc_hsrx *r, const uint8_t *in, size_t n_in, uint8_t *sym_out, size_t *n_sym,
		    float *freq_out, size_t *n_freq) {
  const int chunk_size = 128;
  signed long freq_alpha = 0.04 * 65536;
  signed long freq_beta = 0.0012 * 256 * 65536 / r->omega * r->pll_adjustment;
  float gain_mu = 0.02 / (75.0f*75.0f) * 2;            /* cstln_amp*cstln_amp is float (sdr.h:287) */
  size_t done = 0, ns = 0, nf = 0;
  while ( n_in - done >= (size_t)chunk_size + 1 ) {
    const uint8_t *pin = in + 2*done, *pend = pin + 2*chunk_size;
    uint8_t s_re = 0, s_im = 0;
    uint16_t symbol_arg = 0;
    while ( pin < pend ) {
      if ( r->mu < 1 ) {
	uint16_t a0 = (uint16_t)(r->polar_a[pin[0]][pin[1]] - r->phase) >> 8;
	const uint8_t *p0r = r->rect[a0][r->polar_r[pin[0]][pin[1]] >> 1];
	uint16_t a1 = (uint16_t)(r->polar_a[pin[2]][pin[3]] - (r->phase + r->freqw)) >> 8;
	const uint8_t *p1r = r->rect[a1][r->polar_r[pin[2]][pin[3]] >> 1];
	s_re = (int)(p0r[0] + (p1r[0] - p0r[0]) * r->mu);
	s_im = (int)(p0r[1] + (p1r[1] - p0r[1]) * r->mu);
	symbol_arg = r->polar_a[s_re][s_im];
	int quadrant = symbol_arg >> 14;
	static const unsigned char quadrant_to_symbol[4] = { 0, 2, 3, 1 };
	sym_out[ns++] = quadrant_to_symbol[quadrant];
	int16_t phase_error = (int16_t)(symbol_arg & 16383) - 8192;
	r->phase += (phase_error * freq_alpha + 32768) >> 16;
	r->freqw += (phase_error * freq_beta + 32768*256) >> 24;
	r->hist[2] = r->hist[1];
	r->hist[1] = r->hist[0];
	r->hist[0].p_re = s_re; r->hist[0].p_im = s_im;
	const uint8_t *cp = r->sincos[(uint16_t)((symbol_arg & 49152) + 8192)];
	r->hist[0].c_re = cp[0]; r->hist[0].c_im = cp[1];
	int muerr =
	  ( (signed char)(r->hist[0].p_re - r->hist[2].p_re) * ((int)r->hist[1].c_re - 128) +
	    (signed char)(r->hist[0].p_im - r->hist[2].p_im) * ((int)r->hist[1].c_im - 128) ) -
	  ( (signed char)(r->hist[0].c_re - r->hist[2].c_re) * ((int)r->hist[1].p_re - 128) +
	    (signed char)(r->hist[0].c_im - r->hist[2].c_im) * ((int)r->hist[1].p_im - 128) );
	float mucorr = muerr * gain_mu;
	const float max_mucorr = 0.1;
	if ( mucorr < -max_mucorr ) mucorr = -max_mucorr;
	if ( mucorr >  max_mucorr ) mucorr =  max_mucorr;
	r->mu += mucorr;
	r->mu += r->omega;
      }
      pin += 2;
      --r->mu;
      r->phase += r->freqw;
    }
    done += chunk_size;
    if ( !r->allow_drift ) {                           /* sdr.h:1122-1125 */
      if ( r->freqw < r->min_freqw || r->freqw > r->max_freqw )
	r->freqw = (r->max_freqw + r->min_freqw) / 2;
    }
    r->meas_count += chunk_size;                       /* sdr.h:1129-1134 */
    while ( r->meas_count >= r->meas_decimation ) {
      r->meas_count -= r->meas_decimation;
      if ( freq_out ) freq_out[nf] = (float)r->freqw / 65536;
      ++nf;
    }
  }
  *n_sym = ns;
  if ( n_freq ) *n_freq = nf;
  return done;
}

/* ----------------------------------------------------- dvb_deconvol_sync_hard */

void orc_hsdeconv_init(orc_hsdeconv *d, int resync_period) {    /* dvb.h:622-631, 674-705 */
  static const uint8_t luts[4][4] = { {0,1,2,3}, {2,0,3,1}, {1,0,3,2}, {0,2,1,3} };
  memset(d, 0, sizeof *d);
  d->resync_period = resync_period;
  memcpy(d->lut, luts, sizeof luts);
}

/* deconvol_poly2<u8,uint32_t,uint64_t,0x3ba,0x38f70>::run (convolutional.h:96-187), nb = 64. */
static int hs_poly2_run(uint32_t *inI, uint32_t *inQ, const uint8_t *pin, const uint8_t *remap,
			uint8_t *pout, int nb) {
  const uint64_t POLY_DECONVOL = 0x3ba, POLY_ERRORS = 0x38f70;
  nb /= 4;
  unsigned long nerrors = 0;
  int halfway = nb / 2;
  uint32_t histI = *inI, histQ = *inQ;
  for ( ; nb--; ) {
    uint32_t wd = 0, we = 0;
    for ( int bit = 32; bit--; ++pin ) {
      uint8_t iq = remap[*pin];
      histI = (histI << 1) | (iq >> 1);
      histQ = (histQ << 1) | (iq & 1);
      if ( POLY_DECONVOL & ((uint64_t)2 << (2*bit)) ) wd ^= histI;
      if ( POLY_DECONVOL & ((uint64_t)1 << (2*bit)) ) wd ^= histQ;
      if ( POLY_ERRORS   & ((uint64_t)2 << (2*bit)) ) we ^= histI;
      if ( POLY_ERRORS   & ((uint64_t)1 << (2*bit)) ) we ^= histQ;
    }
    *pout++ = wd >> 24; *pout++ = wd >> 16; *pout++ = wd >> 8; *pout++ = wd;
    if ( nb < halfway ) nerrors += __builtin_popcount(we);
  }
  *inI = histI; *inQ = histQ;
  return nerrors;
}

/* dvb_deconvol_sync::run (dvb.h:633-660): 512 symbols -> 64 bytes per iteration. */
size_t orc_hsdeconv_run(orc_hsdeconv *d, const uint8_t *sym, size_t n_in, uint8_t *out, size_t out_cap,
			size_t *consumed) {
  const int chunk_size = 64;
  size_t rd = 0, wr = 0;
  while ( n_in - rd >= (size_t)chunk_size*8 && out_cap - wr >= (size_t)chunk_size ) {
    int errors_best = 1 << 30;
    int best = -1;
    for ( int s = 0; s < 4; ++s ) {
      if ( d->resync_phase != 0 && s != d->locked ) continue;
      uint8_t dummy[64];
      uint8_t *pout = (s == d->locked) ? out + wr : dummy;
      int nerrors = hs_poly2_run(&d->inI[s], &d->inQ[s], sym + rd, d->lut[s], pout, chunk_size);
      if ( nerrors < errors_best ) { errors_best = nerrors; best = s; }
    }
    rd += chunk_size*8;
    wr += chunk_size;
    if ( best != d->locked ) d->locked = best;
    if ( ++d->resync_phase >= d->resync_period ) d->resync_phase = 0;
  }
  *consumed = rd;
  return wr;
}
