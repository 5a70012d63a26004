/* leandvb_b200_tx.h -- C ABI of the B200-native leandvbtx DVB-S transmit chain
 * (SURVEY.md section 8, rows a5 and "next" 2: it synthesises IQ in HBM, so that
 * the receive path can be measured without the PCIe ceiling, and it is what a
 * `runnable` inside the reference's scheduler binds to in apps/leandvbtx.cc).
 *
 * One handle replaces the chain of reference runnables between p_tspackets and
 * the output file_writer in apps/leandvbtx.cc:79-197:
 *
 *   randomizer          leansdr/dvb.h:1063-1102
 *   rs_encoder          leansdr/dvb.h:957-980, rs.h:141-167
 *   interleaver         leansdr/dvb.h:900-921
 *   dvb_convol          leansdr/dvb.h:519-604, convolutional.h:225-270
 *   cstln_transmitter   leansdr/sdr.h:1196-1221
 *   fir_resampler       leansdr/dsp.h:290-364   (RRC interpolation by `interp`)
 *   decimator           leansdr/generic.h:247-267
 *   simple_agc          leansdr/sdr.h:238-274   (--agc)
 *
 * The output is bit-identical to the reference's cf32 stream: the taps are built
 * on the host with the reference's expressions (leandvbtx.cc:131-138), every
 * product and sum on the device is a separate round-to-nearest operation in the
 * reference's order.  Error convention and threading as in leandvb_b200.h.
 */
#ifndef LEANDVB_B200_TX_H
#define LEANDVB_B200_TX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors leandvbtx's `struct config` (apps/leandvbtx.cc:56-77), same defaults
 * (ldvbtx_config_default).  constellation / fec use the LDVB_CSTLN_* / LDVB_FEC*
 * values of leandvb_b200.h. */
typedef struct ldvbtx_config {
  uint32_t abi_version;      /* = LDVB_ABI_VERSION                              */
  int32_t  constellation;    /* --const   BPSK | QPSK | 8PSK                    */
  int32_t  fec;              /* --cr      (2/3 on QPSK is handled as 4/6, leandvbtx.cc:115-119) */
  int32_t  interp, decim;    /* -f INTERP[/DECIM]                               */
  float    rolloff;          /* --roll-off                                      */
  float    rrc_rej;          /* --rrc-rej                                       */
  char     power_db[32];     /* --power, as typed: amp = expf(logf(10)*atof(s)/20) (leandvbtx.cc:289) */
  int32_t  agc;              /* --agc                                           */
  int32_t  device;           /* CUDA device ordinal                             */
  uint64_t max_packets;      /* largest n_packets passed to one call            */
  int32_t  keep_taps;        /* keep intermediate streams for ldvbtx_tap()      */
  int32_t  reserved[3];
} ldvbtx_config;

typedef struct ldvbtx_handle ldvbtx_handle;

enum {
  LDVBTX_TAP_RSPACKETS = 0,  /* 204-byte packets after randomizer + rs_encoder (p_rspackets)  */
  LDVBTX_TAP_MPEGBYTES = 1,  /* interleaved bytes (p_mpegbytes)                                */
  LDVBTX_TAP_SYMBOLS   = 2   /* u8 constellation symbols (p_symbols)                           */
};

void ldvbtx_config_default(ldvbtx_config *cfg);              /* leandvbtx.cc:69-76 */
int  ldvbtx_create(const ldvbtx_config *cfg, ldvbtx_handle **out);
int  ldvbtx_destroy(ldvbtx_handle *h);
int  ldvbtx_reset(ldvbtx_handle *h);                          /* back to the start of a stream */
const char *ldvbtx_last_error(const ldvbtx_handle *h);
int  ldvbtx_set_stream(ldvbtx_handle *h, void *cuda_stream);

/* Upper bound of the samples one call with n_packets can return (for sizing buffers). */
size_t ldvbtx_max_samples(const ldvbtx_handle *h, size_t n_packets);

/* Host in, host out: what a runnable's run() calls.  n_packets 188-byte TS packets in,
 * *n_samples cf32 samples (interleaved I,Q floats) out.  Items a stage cannot consume yet
 * (the interleaver's 11 packets, the filter's history, a partial AGC chunk) stay in the
 * handle, like unread items in the reference's pipebufs. */
int ldvbtx_push(ldvbtx_handle *h, const uint8_t *ts_host, size_t n_packets,
		float *iq_host, size_t cap_samples, size_t *n_samples);

/* Same with device pointers on cfg.device: TS packets in HBM, IQ left in HBM. */
int ldvbtx_process_device(ldvbtx_handle *h, const uint8_t *ts_dev, size_t n_packets,
			  float *iq_dev, size_t cap_samples, size_t *n_samples);

/* leantsgen's numbered packets (apps/leantsgen.cc:37-47) written straight into HBM:
 * packets first .. first+n-1. */
int ldvbtx_tsgen_device(ldvbtx_handle *h, uint64_t first, size_t n_packets, uint8_t *ts_dev);

/* Intermediate stream produced by the LAST call (keep_taps != 0). */
int ldvbtx_tap(ldvbtx_handle *h, int which, void *dst_host, size_t cap_bytes, size_t *n_bytes);
/* The interpolation taps in use (ncoeffs floats). */
int ldvbtx_taps(ldvbtx_handle *h, float *dst_host, size_t cap_floats, size_t *n_floats);
/* Same from a configuration alone: pure host code, needs no device. */
int ldvbtx_host_taps(const ldvbtx_config *cfg, float *dst_host, size_t cap_floats, size_t *n_floats);

/* Stand-alone fir_resampler<cf32,float> (dsp.h:290-364) with decim = 1:
 *   y[n*interp + p] = sum_j taps[p + j*interp] * x[n + latency - j],  latency = (ntaps+interp)/interp,
 * complex taps (re,im) = shifted_coeffs (dsp.h:352-361), accumulated in that order from 0.
 * Host in, host out; *n_out = ((n_in*interp - ntaps)/interp)*interp (0 while n_in < ntaps). */
int ldvbtx_fir_resampler_cf32(int device, const float *x_host, size_t n_in,
			      const float *taps_cplx, uint32_t ntaps, uint32_t interp,
			      float *y_host, size_t cap_out, size_t *n_out);

#ifdef __cplusplus
}
#endif
#endif /* LEANDVB_B200_TX_H */
