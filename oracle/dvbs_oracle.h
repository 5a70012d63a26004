/* oracle/dvbs_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, scalar, one thread) of the leandvb DVB-S receive
 * path of pabr/leansdr, written from the behaviour of the reference sources
 * (cited per function as /root/reference/src/<file>:<lines>).  It is the
 * checker for the CUDA product path: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load it.  The product (leansdr_b200/) never
 * includes, links or calls anything in this directory.
 *
 * Parity pin: every stage below is compared bit-for-bit with streams tapped
 * from the UNMODIFIED reference runnables (oracle/ref_tap.cc, built into
 * oracle/_ref/ from the sources where they lie) in tests/test_oracle_vs_ref.py
 * and with the committed golden vectors under tests/golden/.
 */
#ifndef DVBS_ORACLE_H
#define DVBS_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- tables */

enum { ORC_BPSK = 0, ORC_QPSK = 1, ORC_8PSK = 2, ORC_16APSK = 3, ORC_32APSK = 4, ORC_64APSKE = 5,
       ORC_16QAM = 6, ORC_64QAM = 7, ORC_256QAM = 8 };
enum { ORC_FEC12 = 0, ORC_FEC23 = 1, ORC_FEC46 = 2, ORC_FEC34 = 3,
       ORC_FEC56 = 4, ORC_FEC78 = 5 };

/* One cell of the 256x256 constellation table (sdr.h:455-458, 529-560). */
typedef struct {
  int16_t cost;
  int16_t symbol;        /* 0..nsymbols-1 */
  int16_t phase_error;
  int16_t pad;
} orc_cstln_cell;

typedef struct {
  orc_cstln_cell cell[256][256];  /* indexed [(u8)I][(u8)Q] */
  int8_t sym_re[256], sym_im[256];
  int nsymbols, nrotations;
} orc_cstln;

/* fec (ORC_FEC*) selects the APSK ring ratios (dvb.h:45-81); returns 0, or -1 where the
   reference fail()s ("Code rate not supported with APSK16/32"). */
int orc_cstln_build2(orc_cstln *c, int kind, int fec, int harden);
void orc_cstln_build(orc_cstln *c, int kind, int harden);   /* fec = 3/4 for the APSKs */
void orc_trig16_build(float *lut /* [65536][2] = cos,sin */);
void orc_rs_tables(uint8_t *exp511, uint8_t *log256, uint8_t *gen17);
void orc_derand_pattern(uint8_t *pat1504);
/* filtergen.h:45-62 followed by normalize_dcgain (leandvb.cc:375-376). */
int  orc_lowpass(int order, float Fcut, float *coeffs /* [order+1] */);
int  orc_resample_design(float Fs, float Fm, float rolloff, float rej,
			 unsigned decim_opt, float *coeffs, int max_coeffs,
			 int *decim_out);
/* filtergen.h:68-92. Returns ncoeffs. */
int  orc_rrc(int order, float Fs, float rolloff, float *coeffs);
/* dvb.h:205-292: deconvolution polynomials. Returns punctperiod. */
int  orc_deconv_polys(int fec, uint64_t *deconv, uint64_t *deconv2,
		      int *punctweight);

/* ----------------------------------------------------------- front end */

/* dsp.h:33-54 instantiated <u8,128,f32,0,1,1> etc. fmt: 0=u8 1=s8 2=u16 3=s16 */
void orc_cconvert(const void *in, int fmt, float *out_cf32, size_t n);
/* dsp.h:140-160 */
void orc_scale(const float *in_cf32, float scale, float *out_cf32, size_t n);

/* sdr.h:1228-1261 */
typedef struct {
  float lut_cos[65536], lut_sin[65536];
  uint16_t index;
} orc_rotator;
void orc_rotator_init(orc_rotator *r, float freq);
void orc_rotator_run(orc_rotator *r, const float *in, float *out, size_t n);

/* dsp.h:219-285.  shifted = complex taps [ncoeffs][2]. */
typedef struct {
  unsigned ncoeffs, decim;
  const float *coeffs;
  float *shifted;
  float current_freq;
} orc_fir;
void   orc_fir_init(orc_fir *f, unsigned ncoeffs, const float *coeffs, unsigned decim);
void   orc_fir_set_freq(orc_fir *f, float freq);
/* Consumes as the reference run() does given n_in readable samples and
 * unlimited output space.  Returns outputs written, *consumed = samples read. */
size_t orc_fir_run(orc_fir *f, const float *in, size_t n_in, float *out,
		   size_t *consumed);

/* generic.h:247-267 */
size_t orc_decimate(const float *in, size_t n_in, unsigned d, float *out,
		    size_t *consumed);

/* sdr.h:46-154 + dsp.h:56-116 */
#define ORC_NOTCH_N 4096
#define ORC_NOTCH_MAXSLOTS 8
typedef struct {
  int nslots;
  int phase;
  float gain, k;
  int decimation;
  float agc_rms_setpoint;
  struct {
    int i;
    float estim_re, estim_im;
    float *expj;            /* [4096][2] */
  } slots[ORC_NOTCH_MAXSLOTS];
  /* fft tables */
  int *bitrev;
  float *omega_rev;         /* [4096][2] */
} orc_notch;
void   orc_notch_init(orc_notch *a, int nslots);
/* processes whole 4096-blocks; returns samples consumed (= produced). */
size_t orc_notch_run(orc_notch *a, const float *in, size_t n_in, float *out);
/* cfft_engine::inplace (dsp.h:78-110) exposed for tests. */
void   orc_fft_inplace(int n, float *data, int reverse);

/* ------------------------------------------------------------ cnr_fft / spectrum
 * sdr.h:1273-1345 (cnr_fft<f32>, nfft 4096) and sdr.h:1347-1404 (spectrum<f32>, nfft 1024).
 * Both read the stream in front of the FIR (leandvb.cc:322-343).  One struct serves both:
 * n = 4096 + bandwidth > 0 -> CNR values; n = 1024 + bandwidth = 0 -> 1024-bin rows. */
typedef struct {
  int n;
  float bandwidth;          /* Fm/Fs (cnr) */
  float kavg;               /* 0.1 (cnr), 0.5 (spectrum as leandvb sets it, :342) */
  int decimation;           /* decimation(Fs, 1), leandvb.cc:328,341 */
  int phase;
  int have_avg;
  float avgpower[4096];
} orc_meas;
void   orc_meas_init(orc_meas *m, int n, float bandwidth, float kavg, int decimation);
/* Consumes whole n-sample blocks of cf32 `in`; center_freq = *freq_tap * tap_multiplier at
 * the time of the call (cnr only).  Writes one float per CNR point, or n floats per spectrum
 * row, up to cap points; returns the number of points, *consumed = samples read. */
size_t orc_meas_run(orc_meas *m, const float *in, size_t n_in, float center_freq, float *out,
		    size_t cap_points, size_t *consumed);

/* ------------------------------------------------------------ receiver */

enum { ORC_SAMP_NEAREST = 0, ORC_SAMP_LINEAR = 1, ORC_SAMP_RRC = 2 };

typedef struct {
  /* configuration */
  const orc_cstln *cstln;
  const float *trig;        /* [65536][2] */
  int sampler;
  unsigned long meas_decimation;
  float omega, min_omega, max_omega;
  float freqw, min_freqw, max_freqw;
  float pll_adjustment;
  int allow_drift;
  float kest;
  /* rrc sampler */
  int rrc_ncoeffs, rrc_sub;
  const float *rrc_coeffs;
  float *rrc_shifted;       /* [ncoeffs][2] */
  int rrc_update_phase;
  /* linear sampler copy of freqw (sdr.h:625) */
  float samp_freqw;
  /* state (sdr.h:921-934) */
  float est_insp, agc_gain, mu, phase, est_sp, est_ep;
  unsigned long meas_count;
  struct { float p_re, p_im, c_re, c_im; } hist[3];
  float freq_tap;
} orc_rx;

void orc_rx_init(orc_rx *r, const orc_cstln *c, const float *trig, int sampler);
void orc_rx_set_omega(orc_rx *r, float omega);
void orc_rx_set_freq(orc_rx *r, float freq);
void orc_rx_set_rrc(orc_rx *r, int ncoeffs, const float *coeffs, int subsampling);
int  orc_rx_readahead(const orc_rx *r);
/* Runs whole 128-sample chunks while n_in >= 128+readahead (sdr.h:783-915).
 * symbols_out: 4 bytes per symbol {int16 cost, u8 symbol, 0}.
 * sampled_out (optional): one cf32 per chunk that produced a symbol.
 * meas_out (optional): per measurement {freq_tap, ss, mer} floats.
 * Returns samples consumed. */
size_t orc_rx_run(orc_rx *r, const float *in, size_t n_in,
		  uint8_t *symbols_out, size_t *n_symbols,
		  float *sampled_out, size_t *n_sampled,
		  float *meas_out, size_t *n_meas);

/* ----------------------------------------------- deconvolution and sync */

typedef struct {
  uint8_t lut[2][2];
  uint64_t in;  int n_in;
  uint64_t out; int n_out;
  uint64_t in2; int n_in2, n_out2;   /* auxiliary register for fastlock (dvb.h:303-306) */
} orc_dsync;

typedef struct {
  int punctperiod, punctweight;
  uint64_t deconv[8], deconv2[8];
  orc_dsync syncs[4];
  int locked;
  int skip;
  int fastlock;                      /* dvb.h:195, 428-454 */
} orc_deconv;

void orc_deconv_init(orc_deconv *d, int fec);
void orc_deconv_next_sync(orc_deconv *d);          /* dvb.h:185-193 */
/* One reference run() (dvb.h:414-467, fastlock off) with n_in symbols
 * readable and out_cap bytes writable. Returns bytes written. */
size_t orc_deconv_run(orc_deconv *d, const uint8_t *symbols4, size_t n_in,
		      uint8_t *out, size_t out_cap, size_t *consumed);

size_t orc_deconv_run2(orc_deconv *d, const uint8_t *symbols4, size_t n_in,
		       uint8_t *out, size_t out_cap, size_t *consumed, int big_batch);
void   orc_deconv_set(orc_deconv *d, int locked, int skip);
void   orc_deconv_set_fastlock(orc_deconv *d, int on);

typedef struct {
  int scan_syncs, want_syncs;
  unsigned long lock_timeout;
  uint8_t polarity;
  int bitphase;
  int synchronized;
  int next_sync_count;
  int phase8;
  unsigned long lock_timeleft, locktime;
  int report_state;
  int fastlock, resync_period, resync_phase;   /* dvb.h:716-717, 781-796 */
} orc_mpegsync;

void orc_mpegsync_init(orc_mpegsync *m);
void orc_mpegsync_set_fastlock(orc_mpegsync *m, int fastlock, int resync_period);
/* One reference run() (dvb.h:742-874, fastlock off).  deconv may be NULL.
 * lock_out receives lock transitions (0/1), locktime_out one per packet. */
size_t orc_mpegsync_run(orc_mpegsync *m, orc_deconv *deconv,
			const uint8_t *in, size_t n_in,
			uint8_t *out, size_t out_cap, size_t *consumed,
			int *lock_out, size_t *n_lock,
			uint64_t *locktime_out, size_t *n_locktime);

size_t orc_mpegsync_run2(orc_mpegsync *m, orc_deconv *deconv,
			 const uint8_t *in, size_t n_in,
			 uint8_t *out, size_t out_cap, size_t *consumed,
			 int *lock_out, size_t *n_lock,
			 uint64_t *locktime_out, size_t *n_locktime,
			 int per_wrap, int *switched);

/* dvb.h:926-948. Returns packets written. */
size_t orc_deinterleave(const uint8_t *in, size_t n_in, uint8_t *out_packets,
			size_t *consumed);

/* dvb.h:985-1058 + rs.h:86-268.  One packet. Returns 1 if still corrupted. */
int orc_rs_decode_packet(uint8_t *pin204 /* fixed in place like rs.h:261 */,
			 uint8_t *pout188, int *bits_corrected);
/* rs.h:131-160 (for round-trip tests) */
void orc_rs_encode(uint8_t *msg204);

typedef struct { int pos; } orc_derand;
/* dvb.h:1107-1163.  Returns packets written (dropped packets are skipped). */
size_t orc_derandomize(orc_derand *d, const uint8_t *in188, size_t npackets,
		       uint8_t *out188);

/* ------------------------------------------------------------- Viterbi */

typedef struct orc_viterbi orc_viterbi;
orc_viterbi *orc_viterbi_new(const orc_cstln *c, int fec);
void   orc_viterbi_free(orc_viterbi *v);
void   orc_viterbi_set_resync_period(orc_viterbi *v, int p);
int    orc_viterbi_nsyncs(const orc_viterbi *v);
int    orc_viterbi_current_sync(const orc_viterbi *v);
/* dvb.h:1353-1414. Returns bytes written. */
/* model-checking hooks for the time-segment schedule of the CUDA Viterbi stage */
size_t orc_viterbi_chunk(orc_viterbi *v, const uint8_t *symbols4, uint32_t run_mask, int out_sync,
			 uint8_t *out, int32_t *totals);
void   orc_viterbi_get_dec(const orc_viterbi *v, int s, int32_t *cost64, uint64_t *path64);
void   orc_viterbi_set_dec(orc_viterbi *v, int s, const int32_t *cost64, const uint64_t *path64);
void   orc_viterbi_set_ctl(orc_viterbi *v, int current_sync, int resync_phase);
int    orc_viterbi_resync_phase(const orc_viterbi *v);
int    orc_viterbi_nshifts(const orc_viterbi *v);
int    orc_viterbi_bits_in(const orc_viterbi *v);
size_t orc_viterbi_run(orc_viterbi *v, const uint8_t *symbols4, size_t n_in,
		       uint8_t *out, size_t out_cap, size_t *consumed);

/* ------------------------------------------------------------ transmit chain
 * oracle/dvbs_tx_oracle.c: leandvbtx (apps/leandvbtx.cc:79-197), stage by stage. */
void   orc_tx_randomize(const uint8_t *ts, size_t npk, uint8_t *out);
void   orc_tx_rs_encode(const uint8_t *ts188, size_t npk, uint8_t *rs204);
size_t orc_tx_interleave(const uint8_t *rs204, size_t npk, uint8_t *out);
size_t orc_tx_convol(int fec, int bps, const uint8_t *in, size_t n, uint8_t *out, size_t *consumed);
void   orc_tx_map(const orc_cstln *c, const uint8_t *sym, size_t n, float *out_cf32);
float  orc_tx_amp(const char *power_db);
int    orc_tx_taps(int interp, float rolloff, float rrc_rej, float amp, float *coeffs);
size_t orc_tx_resample(const float *xin, size_t n_in, const float *coeffs, int ncoeffs, int interp,
		       float *yout, size_t *consumed);
size_t orc_tx_agc(const float *xin, size_t n, float out_rms, float bw, float *yout);
size_t orc_tx_chain(const uint8_t *ts, size_t npk, int cstln_kind, int fec, int interp, int decim,
		    float rolloff, float rrc_rej, const char *power_db, int agc,
		    float *out, size_t cap_samples,
		    uint8_t *tap_mpegbytes, size_t *n_mpegbytes, uint8_t *tap_symbols, size_t *n_symbols);

/* ------------------------------------------------------------- --hs path
 * oracle/dvbs_hs_oracle.c: fast_qpsk_receiver<u8> (sdr.h:946-1189) and
 * dvb_deconvol_sync_hard (dvb.h:612-707, convolutional.h:75-192). */
typedef struct {
  unsigned long meas_decimation;
  float omega, min_omega, max_omega;
  signed long freqw, min_freqw, max_freqw;
  float pll_adjustment;
  int allow_drift;
  uint16_t polar_a[256][256]; uint8_t polar_r[256][256];   /* lut_polar {a, r} */
  uint8_t rect[256][256][2];                                 /* lut_rect[angle][r] {re, im} */
  uint8_t sincos[65536][2];                                  /* lut_sincos {re, im} */
  struct { uint8_t p_re, p_im, c_re, c_im; } hist[3];
  float mu;
  uint16_t phase;
  unsigned long meas_count;
} orc_hsrx;
void   orc_hsrx_init(orc_hsrx *r);
void   orc_hsrx_set_omega(orc_hsrx *r, float omega);
void   orc_hsrx_set_freq(orc_hsrx *r, float freq);
void   orc_hsrx_config(orc_hsrx *r, int allow_drift, unsigned long meas_decimation);
size_t orc_hsrx_run(orc_hsrx *r, const uint8_t *in, size_t n_in, uint8_t *sym_out, size_t *n_sym,
		    float *freq_out, size_t *n_freq);

typedef struct {
  int resync_period, resync_phase, locked;
  uint32_t inI[4], inQ[4];
  uint8_t lut[4][4];
} orc_hsdeconv;
void   orc_hsdeconv_init(orc_hsdeconv *d, int resync_period);
size_t orc_hsdeconv_run(orc_hsdeconv *d, const uint8_t *sym, size_t n_in, uint8_t *out, size_t out_cap,
			size_t *consumed);

#ifdef __cplusplus
}
#endif
#endif
