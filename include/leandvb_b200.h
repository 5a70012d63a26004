/* leandvb_b200.h -- C ABI of the B200-native leandvb DVB-S receive path.
 *
 * This is the drop-in boundary: plain C types, pointers and sizes only.  The
 * entry points are what a `runnable` living inside the reference's own
 * scheduler/pipebuf framework binds to (see INTEGRATION.md for the thin
 * `gpu_dvbs_receiver : runnable` wrapper a maintainer adds to leandvb.cc).
 * Citations are file:line under /root/reference/src/.
 *
 * One handle replaces the chain of reference runnables between the input
 * pipebuf and `p_tspackets` in apps/leandvb.cc:204-596:
 *
 *   cconverter / scaler          leansdr/dsp.h:33-54, 140-160
 *   auto_notch                   leansdr/sdr.h:46-154
 *   rotator                      leansdr/sdr.h:1228-1261
 *   fir_filter (+ decimator)     leansdr/dsp.h:219-285, generic.h:247-267
 *   cstln_receiver + samplers    leansdr/sdr.h:589-938
 *   deconvol_sync | viterbi_sync leansdr/dvb.h:122-476 | 1173-1416
 *   mpeg_sync                    leansdr/dvb.h:712-891
 *   deinterleaver                leansdr/dvb.h:926-948
 *   rs_decoder                   leansdr/dvb.h:985-1058, rs.h:86-268
 *   derandomizer                 leansdr/dvb.h:1107-1163
 *   (--hs) fast_qpsk_receiver    leansdr/sdr.h:946-1189
 *   (--hs) dvb_deconvol_sync_hard leansdr/dvb.h:612-707, convolutional.h:75-192
 *
 * Error convention: every function returns 0 on success or a negative
 * LDVB_E* code; ldvb_strerror() gives text.  The runnable wrapper turns a
 * non-zero code into the reference's fail()/exit(1) (framework.h:32-33).
 * Nothing here throws, allocates on behalf of the caller, or takes ownership
 * of caller buffers.  A handle is single-threaded like the reference
 * (README.coding.md:29): one caller at a time.
 */
#ifndef LEANDVB_B200_H
#define LEANDVB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDVB_ABI_VERSION 3

/* ------------------------------------------------------------ error codes */
enum {
  LDVB_OK          = 0,
  LDVB_EINVAL      = -1,   /* bad argument / unsupported configuration   */
  LDVB_ENOMEM      = -2,   /* host or device allocation failed            */
  LDVB_ECUDA       = -3,   /* CUDA runtime error (see ldvb_last_error)    */
  LDVB_ENODEV      = -4,   /* no usable sm_100 device                     */
  LDVB_EOVERFLOW   = -5,   /* more data than the handle was sized for     */
  LDVB_ESTATE      = -6    /* call sequence error                         */
};

/* ------------------------------------------------------------- enumerations
 * Values follow the order of the reference's own enums so that a wrapper can
 * cast: config::input_format (leandvb.cc:46-50), cstln_lut::predef
 * (sdr.h:305-311), code_rate (dvb.h:36-40), config::sampler (leandvb.cc:70). */
enum { LDVB_FMT_U8 = 0, LDVB_FMT_S8 = 1, LDVB_FMT_U16 = 2, LDVB_FMT_S16 = 3,
       LDVB_FMT_F32 = 4 };
/* cstln_lut<256>::predef (sdr.h:305-311).  The APSK ring ratios follow the code rate
 * (make_dvbs2_constellation, dvb.h:45-81): 16APSK needs 2/3, 3/4 or 5/6, 32APSK 3/4 or 5/6;
 * other combinations are LDVB_EINVAL where the reference fail()s. */
enum { LDVB_CSTLN_BPSK = 0, LDVB_CSTLN_QPSK = 1, LDVB_CSTLN_8PSK = 2,
       LDVB_CSTLN_16APSK = 3, LDVB_CSTLN_32APSK = 4, LDVB_CSTLN_64APSKE = 5,
       LDVB_CSTLN_16QAM = 6, LDVB_CSTLN_64QAM = 7, LDVB_CSTLN_256QAM = 8 };
enum { LDVB_FEC12 = 0, LDVB_FEC23 = 1, LDVB_FEC46 = 2, LDVB_FEC34 = 3,
       LDVB_FEC56 = 4, LDVB_FEC78 = 5 };
enum { LDVB_SAMP_NEAREST = 0, LDVB_SAMP_LINEAR = 1, LDVB_SAMP_RRC = 2 };

/* Receiver scheduling mode.
 *   EXACT: the symbol-timing / carrier recurrence (sdr.h:800-847) is walked
 *          serially from the carried state: every softsymbol field is
 *          bit-identical to the reference.
 *   FAST:  the stream is cut into time spans that run concurrently from a
 *          warm-up state; spans are stitched on symbol time, their 90-degree
 *          ambiguity is resolved against the previous span, and every seam is
 *          verified: by default (seam_mode 0) a seam stands only when EVERY hard
 *          decision in the overlap agrees and the loop states on both sides agree
 *          within tight bounds; otherwise the later span is re-run exactly from the
 *          earlier span's end state (at most two repair rounds per batch, after
 *          which the tolerant rule -- <= 1/16 mismatches -- decides; the count of
 *          seams accepted that way is reported in ldvb_meas).  FAST parity is
 *          STATISTICAL, not by construction: the loop state of a span that was
 *          warmed up is close to, not identical with, the serial state, so soft
 *          costs differ by one table cell on ~1 % of the symbols and an isolated
 *          hard decision can differ at low SNR (bench.py measures both against the
 *          EXACT mode and the reference binary: `fast_vs_exact`).
 *          Constant-envelope constellations only (BPSK, QPSK, 8PSK: the slicer
 *          looks at the angle alone).  The ring / grid decisions of the APSK and
 *          QAM constellations depend on the AGC estimate, which remembers ~100
 *          chunks (sdr.h:863-869) -- more than a span's warm-up can reproduce --
 *          so a handle for those constellations runs EXACT whatever is asked.
 *          AGC settling: spans restart from the batch-entry AGC estimate.  On the
 *          first FAST batch of a stream (after create / reset), and whenever the
 *          measured input power is more than a factor 2 away from the carried
 *          estimate (a stream that is not at the nominal level, e.g. after a heavy
 *          decimation; a level jump), the first `settle_chunks` chunks of the batch
 *          are walked serially (exactly, ~14 ms for the default 512 chunks) and the
 *          spans start from the state reached there.  settle_chunks = -1 disables it:
 *          hard decisions of BPSK/QPSK/8PSK do not depend on the estimate, soft costs
 *          then stay ~10 % off the serial ones for many batches. */
enum { LDVB_RX_EXACT = 0, LDVB_RX_FAST = 1 };

/* ------------------------------------------------------------------ config
 * Mirrors the fields of leandvb's `struct config` that reach the hot path
 * (apps/leandvb.cc:43-136) with the same meaning and defaults
 * (ldvb_config_default).  Derived parameters (filter order, decimation,
 * rrc_steps, PLL constants) are computed inside ldvb_create exactly as
 * leandvb.cc:run() derives them (:353-384, :432-462, :476-502). */
typedef struct ldvb_config {
  uint32_t abi_version;      /* = LDVB_ABI_VERSION                           */
  int32_t  input_format;     /* LDVB_FMT_*           --u8/--s8/--u16/--s16/--f32 */
  float    float_scale;      /* --float-scale (f32 input only)               */
  float    Fs;               /* -f   input sample rate, Hz                   */
  float    Fm;               /* --sr symbol rate, Hz                         */
  int32_t  anf;              /* --anf number of notch slots (0 disables)     */
  float    Fderot;           /* --derotate Hz (0 disables the rotator)       */
  int32_t  resample;         /* --resample: low-pass FIR + decimation        */
  float    resample_rej;     /* --resample-rej                               */
  uint32_t decim;            /* --decim (0 = auto when resampling)           */
  int32_t  sampler;          /* LDVB_SAMP_*          --sampler               */
  int32_t  rrc_steps;        /* --rrc-steps (0 = auto)                       */
  float    rrc_rej;          /* --rrc-rej                                    */
  float    rolloff;          /* --roll-off                                   */
  int32_t  constellation;    /* LDVB_CSTLN_*         --const                 */
  int32_t  fec;              /* LDVB_FEC*            --cr                    */
  int32_t  viterbi;          /* --viterbi                                    */
  int32_t  hard_metric;      /* --hard-metric                                */
  int32_t  fastlock;         /* --fastlock (dvb.h:391-454, 781-796; leandvb.cc:540-565) */
  int32_t  allow_drift;      /* --drift                                      */
  float    Ftune;            /* --tune Hz                                    */
  float    Finfo;            /* measurement rate, Hz (leandvb.cc:117, 502)   */
  /* ---- B200-side knobs (no reference counterpart) ---- */
  int32_t  rx_mode;          /* LDVB_RX_EXACT | LDVB_RX_FAST                 */
  int32_t  device;           /* CUDA device ordinal                          */
  uint64_t max_batch;        /* largest n_samples passed to one push/process */
  uint32_t span_chunks;      /* FAST: 128-sample chunks per span (0 = auto)  */
  uint32_t warmup_chunks;    /* FAST: warm-up chunks before a span (0 = auto)*/
  int32_t  keep_taps;        /* keep intermediate streams for ldvb_tap()     */
  int32_t  push_sub_batch;   /* ldvb_push: samples per pipelined sub-batch (0 = 192 MiB of input) */
  int32_t  cnr;              /* --cnr: cnr_fft (sdr.h:1273-1345), needs Fs > 4 Fm */
  int32_t  spectrum;         /* spectrum (sdr.h:1347-1404); leandvb always runs it
                                (leandvb.cc:333-343), ldvb_config_default sets 1 */
  int32_t  vber;             /* rate_estimator on the RS decoder's counts (leandvb.cc:583-587): values
                                are queued for ldvb_pull_vber                                      */
  int32_t  hs;               /* --hs (leandvb.cc:727-969): fast_qpsk_receiver<u8> + dvb_deconvol_sync_hard +
                                mpeg_sync with fastlock; needs u8 input, QPSK, code rate 1/2; no notch,
                                filter, CNR or spectrum blocks exist on that path                  */
  int32_t  vit_segments;     /* --viterbi: target number of concurrent time segments per batch
                                (0 = one full wave of CTAs; 1 = one serial pass, the reference's own schedule)    */
  int32_t  vit_warm_chunks;  /* --viterbi: warm-up of a cold segment, 128-block chunks (0 = 2; -1 = none at
                                all: a test knob, every segment then fails verification and is re-run) */
  /* ---- ABI 2 ---- */
  int32_t  settle_chunks;    /* FAST: chunks of the serial AGC settling pass (0 = 512, -1 = never)  */
  int32_t  seam_mode;        /* FAST: 0 = strict seams (equality in the overlap, else exact re-run),
                                1 = tolerant (<= 1/16 mismatching hard decisions)                  */
  /* ---- ABI 3 ---- */
  int32_t  async_push;       /* 1: ldvb_push returns as soon as the samples have left the caller's buffer;
                                the chain runs on a thread of the handle and the packets reach ldvb_pull
                                when they are done (ldvb_flush waits for them).  Keeps the copy engine busy
                                across push calls.  0 (default): push returns with its packets ready.  */
  int32_t  reserved0;
} ldvb_config;

typedef struct ldvb_handle ldvb_handle;

/* Telemetry: what the reference emits on p_freq/p_ss/p_mer/p_lock/p_locktime/
 * p_vber (leandvb.cc:600-616), sampled at the end of the last batch. */
typedef struct ldvb_meas {
  float    freq_tap;         /* cstln_receiver::freq_tap (sdr.h:917-919)     */
  float    ss;               /* sqrtf(est_insp)        (sdr.h:908-909)       */
  float    mer;              /* 10*log10(est_sp/est_ep) (sdr.h:910-911)      */
  int32_t  lock;             /* mpeg_sync synchronized (dvb.h:825, 867)      */
  uint64_t locktime;         /* packets since lock     (dvb.h:856-858)       */
  uint64_t rs_bits;          /* sum of rs_decoder nbits (dvb.h:1009)         */
  uint64_t rs_errs;          /* sum of corrected bits  (dvb.h:1032, rs.h:259)*/
  uint64_t ts_packets;       /* TS packets emitted so far                    */
  uint64_t ts_dropped;       /* packets dropped by the derandomizer (dvb.h:1146-1156) */
  uint64_t samples_in;       /* IQ samples accepted so far                   */
  uint64_t symbols;          /* soft symbols produced so far                 */
  uint32_t seams_total;      /* FAST: span seams stitched                    */
  uint32_t seams_repaired;   /* FAST: seams that failed verification and were re-run exactly */
  uint32_t notch_repaired;   /* notch segments re-run after a carry mismatch */
  uint32_t kernel_launches;  /* CUDA kernels launched by this handle so far  */
  uint32_t vit_segments;     /* --viterbi: time segments decoded concurrently (cold start + warm-up,
                                entry state verified bit for bit against the predecessor's exit)  */
  uint32_t vit_repaired;     /* --viterbi: segments that had not merged and were re-run exactly  */
  /* ---- ABI 2 ---- */
  uint32_t seams_mismatch_accepted; /* FAST: seams accepted with >= 1 mismatching hard decision in the overlap
                                (tolerant rule; 0 in strict mode unless the repair rounds ran out)  */
  uint32_t settle_passes;    /* FAST: serial AGC settling passes run                              */
  float    seam_max_dphase;  /* FAST: largest |phase| / |freqw| / |mu| difference between the state a span   */
  float    seam_max_dfreqw;  /*       entered its chunks with and its predecessor's end state, over the       */
  float    seam_max_dmu;     /*       verified seams so far (phase units of 2pi/65536, modulo the ambiguity)  */
} ldvb_meas;

/* Intermediate streams readable with ldvb_tap() when keep_taps != 0; names
 * follow the reference pipebufs (leandvb.cc:204-596). */
enum {
  LDVB_TAP_PREPROCESSED = 0, /* cf32, input of cstln_receiver (p_preprocessed) */
  LDVB_TAP_SYMBOLS      = 1, /* softsymbol {int16 cost, u8 symbol, 0} (p_symbols) */
  LDVB_TAP_BYTES        = 2, /* u8 (p_bytes)                                 */
  LDVB_TAP_MPEGBYTES    = 3, /* u8 (p_mpegbytes)                             */
  LDVB_TAP_RSPACKETS    = 4, /* 204-byte packets (p_rspackets)               */
  LDVB_TAP_RTSPACKETS   = 5, /* 188-byte packets (p_rtspackets)              */
  LDVB_TAP_RSFLAGS      = 6, /* per packet int32 {corrupted, bits corrected} */
  LDVB_TAP_SAMPLED      = 7, /* cf32, last symbol of each chunk (p_sampled)  */
  LDVB_TAP_MEAS         = 8  /* float {freq_tap, ss, mer} per measurement    */
};

/* Constant tables, as built on the host for upload (ldvb_table). */
enum {
  LDVB_TABLE_CSTLN   = 0,    /* 65536 x {int16 cost, int16 symbol, int16 phase_error, 0} (sdr.h:529-560) */
  LDVB_TABLE_TRIG16  = 1,    /* 65536 x {cos, sin} float (math.h:95-111)     */
  LDVB_TABLE_RS_EXP  = 2,    /* 512 bytes (rs.h:49-60)                       */
  LDVB_TABLE_RS_LOG  = 3,    /* 256 bytes                                    */
  LDVB_TABLE_DERAND  = 4,    /* 1504 bytes (dvb.h:1116-1129)                 */
  LDVB_TABLE_FIR     = 5,    /* ncoeffs float (filtergen.h:45-62)            */
  LDVB_TABLE_RRC     = 6,    /* ncoeffs float (filtergen.h:68-92)            */
  LDVB_TABLE_DECONV  = 7,    /* punctperiod x uint64 (dvb.h:205-292)         */
  LDVB_TABLE_TRELLIS = 8,    /* 64 x NCS x {pred, us} bytes (viterbi.h:61-92) */
  LDVB_TABLE_VITMAP  = 9,    /* nsyncs x {shift, map[nsymbols]} (dvb.h:1336-1351) */
  /* fast_qpsk_receiver::init_lookup_tables (sdr.h:1144-1164), ldvb_host_table only: */
  LDVB_TABLE_HS_POLAR  = 10, /* 65536 x u32: angle | radius << 16, index (u8)re * 256 + (u8)im */
  LDVB_TABLE_HS_RECT   = 11, /* 65536 x u16: re | im << 8, index angle8 * 256 + radius     */
  LDVB_TABLE_HS_SINCOS = 12, /* 65536 x u16: re | im << 8, index angle16                    */
  LDVB_TABLE_FIR_SHIFTED = 13 /* ncoeffs x {re, im} float: the low-pass taps as fir_filter::set_freq shifts them
                                 for the first batch (dsp.h:236-244, 270-280), ldvb_host_table only */
};

/* --------------------------------------------------------------- lifecycle */
void        ldvb_config_default(ldvb_config *cfg);  /* leandvb.cc:88-135 defaults */
int         ldvb_create(const ldvb_config *cfg, ldvb_handle **out);
int         ldvb_destroy(ldvb_handle *h);
const char *ldvb_strerror(int code);
const char *ldvb_last_error(const ldvb_handle *h);  /* detail of the last failure */
int         ldvb_abi_version(void);

/* ------------------------------------------------------------ host-side I/O
 * What the runnable's run() calls.  ldvb_push copies n_samples IQ samples
 * (interleaved I,Q in cfg.input_format) from host memory to the device and
 * runs the whole chain on them; TS packets become available to ldvb_pull in
 * order.  Data that a stage cannot consume yet (partial 128-sample chunk,
 * partial packet, de-interleaver history ...) is carried to the next push,
 * like unread items staying in a reference pipebuf.  ldvb_pull never blocks:
 * it returns up to cap_packets 188-byte packets. */
int ldvb_push(ldvb_handle *h, const void *iq_host, size_t n_samples);
int ldvb_pull(ldvb_handle *h, uint8_t *ts_host, size_t cap_packets,
	      size_t *n_packets);

/* Threads: a handle has ONE caller at a time, with one exception made for streaming hosts: on an async_push
 * handle ldvb_pull may be called from a second thread while another one is inside ldvb_push / ldvb_flush (a
 * producer and a consumer); ldvb_pull only touches the packet queue, under its own lock. */

/* Waits until everything pushed so far has been processed (async_push) and
 * reports an error of the background chain, if any.  A no-op otherwise.  Every
 * entry point except ldvb_push / ldvb_pull does this implicitly. */
int ldvb_flush(ldvb_handle *h);

/* Page-locks a host buffer the caller owns, once, so that every later
 * ldvb_push from inside it is a direct DMA transfer.  Meant for the reference's
 * pipebuf: framework.h:133-146 allocates it once with `new T[size]` (pageable
 * memory) and hands out pointers into it for the life of the process
 * (pipereader::rd(), framework.h:219-221); a pageable source makes the driver
 * stage every copy through its own bounce buffer, synchronously.  ldvb_push
 * accepts pageable memory too (it is only slower).  Returns LDVB_ECUDA when the
 * range cannot be locked (the caller may carry on unregistered). */
int ldvb_host_register(void *ptr, size_t bytes);
int ldvb_host_unregister(void *ptr);

/* ------------------------------------------------------- device-resident I/O
 * Same processing with the IQ batch already in HBM and the TS packets left in
 * HBM: iq_dev and ts_dev are device pointers on cfg.device.  *n_packets is
 * written on return (one small D2H read per call). */
int ldvb_process_device(ldvb_handle *h, const void *iq_dev, size_t n_samples,
			uint8_t *ts_dev, size_t cap_packets,
			size_t *n_packets);

/* Back to the state of a freshly created handle (streams empty, loops at
 * their initial values: what restarting leandvb does), keeping tables and
 * device buffers. */
int ldvb_reset(ldvb_handle *h);

/* Use the caller's CUDA stream (a cudaStream_t passed as void*) for every
 * kernel and copy of this handle, so that the caller's events bracket them. */
int ldvb_set_stream(ldvb_handle *h, void *cuda_stream);

/* Per-kernel device time, measured with CUDA events on the handle's stream
 * around every launch while enabled (bench.py's roofline numbers). */
typedef struct ldvb_kernel_stat {
  char     name[32];
  uint32_t launches;
  float    ms_total;
} ldvb_kernel_stat;
int ldvb_profile(ldvb_handle *h, int enable);     /* enabling clears the counters */
int ldvb_get_profile(ldvb_handle *h, ldvb_kernel_stat *stats, int cap, int *n);

/* ----------------------------------------------------------- introspection */
int ldvb_get_meas(ldvb_handle *h, ldvb_meas *m);
/* Copies the tap stream produced by the LAST push/process into host memory;
 * *n_bytes receives its size (call with dst == NULL to query). */
int ldvb_tap(ldvb_handle *h, int which, void *dst_host, size_t cap_bytes,
	     size_t *n_bytes);
int ldvb_table(ldvb_handle *h, int which, void *dst_host, size_t cap_bytes,
	       size_t *n_bytes);
/* Same tables from a configuration alone: pure host code, needs no device
 * (used by the CPU-only tests to pin the table builders to the reference). */
int ldvb_host_table(const ldvb_config *cfg, int which, void *dst_host,
		    size_t cap_bytes, size_t *n_bytes);

/* Carry state of the serial stages: what rank r hands
 * to rank r+1 in a time-sharded run, and what tests use to compare with the
 * reference's private members.  Opaque blob of ldvb_state_size() bytes.
 * PARTIAL by design: the notch / rotator / receiver loop / deconvolver /
 * mpeg_sync / derandomizer registers and the fir_filter retune.  It does NOT
 * hold the Viterbi decoders, the --hs carry or the unread stream remainders
 * (use it between batches of a drained handle); the complete hand-over between
 * handles is the EDGE blob of ldvb_shard_back(). */
size_t ldvb_state_size(const ldvb_handle *h);
int    ldvb_get_state(ldvb_handle *h, void *blob, size_t cap);
int    ldvb_set_state(ldvb_handle *h, const void *blob, size_t size);

/* Receiver part of the carry state in a documented layout, 22 x uint32
 * (floats bit-cast): mu phase freqw est_insp agc_gain est_sp est_ep
 * hist[3]{p.re,p.im,c.re,c.im} sampler_freqw freq_tap meas_count
 * (sdr.h:921-934). */
int ldvb_get_rx_state(ldvb_handle *h, uint32_t w[22]);
int ldvb_set_rx_state(ldvb_handle *h, const uint32_t w[22]);

/* ---------------------------------------------------- CNR / spectrum telemetry
 * What the reference writes to p_cnr (one float per second of signal, leandvb.cc:322-329)
 * and p_spectrum (float[1024] per second, leandvb.cc:333-343), queued per batch.  The centre
 * bin of cnr_fft follows freq_tap as sampled when the batch starts (the reference samples it
 * whenever its scheduler happens to run the block). */
int ldvb_pull_cnr(ldvb_handle *h, float *dst, size_t cap, size_t *n);
int ldvb_pull_spectrum(ldvb_handle *h, float *dst, size_t cap_rows, size_t *n_rows);
/* What the reference writes to p_vber: rate_estimator<float> (generic.h:272-305) over rs_decoder's
 * (bits corrected, bits processed) counts with sample_size = max(Fm/2, 50000) (leandvb.cc:583-587).
 * The reference adds the counts of one rs_decoder::run() call at a time (1..4 packets with its
 * default buffers); here the threshold is tested after every packet.  Needs cfg.vber. */
int ldvb_pull_vber(ldvb_handle *h, float *dst, size_t cap, size_t *n);

/* ------------------------------------------------------------ time sharding
 * one stream, N handles (one per GPU, one process each).  The
 * stream is cut into consecutive time chunks; chunk k goes to rank k mod N.
 * A rank holds [halo | chunk] in HBM, the halo being the last samples of the
 * previous chunk (the "chunk-edge samples" it received from its neighbour).
 *
 *   ldvb_shard_detect  auto_notch::detect() on the detect points of the chunk
 *                      (needs only raw samples): turns the notch bins in force
 *                      at the chunk start into those at its end, so that the
 *                      neighbour can start before this rank has finished.
 *   ldvb_shard_front   notch + front end + receiver on the chunk, speculatively:
 *                      every segment/span starts from a warm-up inside the halo.
 *                      This is the expensive part and needs nothing from the
 *                      previous rank except the halo and the notch bins.
 *   ldvb_shard_back    receives the previous rank's EDGE (its last span's seam log
 *                      and loop state, deconvolver registers, sync state, unread
 *                      symbols/bytes, de-interleaver history, PRBS position: a few
 *                      KB), stitches and verifies the seam, runs the exact FEC back
 *                      end and produces this rank's EDGE for the next one.
 * Constraints: rx_mode = LDVB_RX_FAST; n_halo, n_chunk and abs_raw0 (absolute
 * index of the first halo sample) multiples of lcm(4096, 128*decimation);
 * n_halo >= ldvb_shard_min_halo() except for the first chunk of the stream
 * (abs_raw0 = 0, n_halo = 0, edge_in = NULL), whose front stage also resets the handle. */
typedef struct ldvb_shard {
  const void *iq_dev;        /* device: [n_halo | n_chunk] samples in cfg.input_format   */
  uint64_t abs_raw0;         /* absolute stream index of iq_dev[0]                       */
  uint64_t n_halo, n_chunk;  /* samples                                                  */
  uint64_t n_halo_next;      /* halo the NEXT chunk will be given (0 for the last chunk) */
  int32_t  last;             /* final chunk of the stream: nothing is left for a successor */
  int32_t  reserved;
  int32_t  bins_before[4];   /* auto_notch slot bins in force at abs_raw0 (-1 = none)    */
  int32_t  bins_after[4];    /* out (ldvb_shard_detect): bins in force where the next
                                chunk's halo starts = its bins_before                    */
} ldvb_shard;

size_t ldvb_edge_size(void);
size_t ldvb_shard_min_halo(const ldvb_handle *h);
int ldvb_shard_detect(ldvb_handle *h, ldvb_shard *s);
int ldvb_shard_front(ldvb_handle *h, const ldvb_shard *s);
/* edge_in: ldvb_edge_size() bytes from the previous chunk's ldvb_shard_back (NULL for the
 * first chunk); edge_out: ldvb_edge_size() bytes for the next chunk (may be NULL). Host memory. */
int ldvb_shard_back(ldvb_handle *h, const void *edge_in, uint8_t *ts_dev, size_t cap_packets,
		    size_t *n_packets, void *edge_out);

/* The ring around those three calls, inside the library: NCCL point-to-point between neighbouring
 * ranks (ncclSend / ncclRecv on two communicators: halo + notch bins, and the EDGE), one process per
 * GPU.  libnccl.so.2 is resolved at run time; LDVB_ENODEV when it is absent.
 *   ldvb_ring_unique_id  one rank creates two ids (ncclGetUniqueId) and hands their bytes to the
 *                        others by whatever means the host has (MPI, a file, torch.distributed ...)
 *   ldvb_ring_init       collective: every rank, same two ids
 *   ldvb_ring_round      this rank's turn for chunk s (s->iq_dev = [halo | chunk] in HBM, the halo
 *                        region is filled by the call; bins_before / bins_after are filled too):
 *                        receive halo + bins, detect, send halo + bins, front, receive EDGE, back,
 *                        send EDGE.  Chunks must be presented in stream order, chunk k on rank k mod N.
 *                        The caller may refill the chunk buffer after the NEXT call has started or
 *                        after ldvb_ring_flush (the outgoing halo is read from it asynchronously).
 *   ldvb_ring_stats      host wall clock per phase {early, front, wait_edge, back}, ms, accumulated */
#define LDVB_RING_ID_BYTES 128
int ldvb_ring_unique_id(void *id, size_t cap_bytes);
int ldvb_ring_init(ldvb_handle *h, const void *id_early, const void *id_edge, int rank, int nranks);
int ldvb_ring_round(ldvb_handle *h, ldvb_shard *s, uint8_t *ts_dev, size_t cap_packets, size_t *n_packets);
int ldvb_ring_flush(ldvb_handle *h);
int ldvb_ring_stats(ldvb_handle *h, double ms4[4], int reset);
int ldvb_ring_destroy(ldvb_handle *h);

/* ------------------------------------------------- stand-alone stage kernels
 * Host in, host out; used by the parity tests and by callers that only need
 * one block.  Each runs the same kernel the chain uses. */

/* fir_filter<cf32,float> with decimation (dsp.h:246-259):
 * y[k] = sum_i taps[i] * x[k*decim + ntaps - i], accumulated in that order.
 * taps are complex (re,im) pairs: the caller passes shifted_coeffs
 * (dsp.h:270-280).  Returns the number of outputs in *n_out. */
int ldvb_fir_cf32(int device, const float *x_host, size_t n_in,
		  const float *taps_cplx, uint32_t ntaps, uint32_t decim,
		  float *y_host, size_t cap_out, size_t *n_out);

/* deinterleaver + rs_decoder on aligned bytes (dvb.h:926-948, 985-1058):
 * mpegbytes -> RS(204,188)-decoded, still randomised packets. */
int ldvb_deint_rs(int device, const uint8_t *mpegbytes_host, size_t n_bytes,
		  uint8_t *rts_host, size_t cap_packets, size_t *n_packets,
		  int32_t *flags_host /* [n][2] corrupted, bits corrected; may be NULL */);

/* rs_decoder alone on n 204-byte packets (dvb.h:1004-1047). */
int ldvb_rs_decode(int device, const uint8_t *rs204_host, size_t n_packets,
		   uint8_t *ts188_host, int32_t *flags_host);

#ifdef __cplusplus
}
#endif
#endif /* LEANDVB_B200_H */
