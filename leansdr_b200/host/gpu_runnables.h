// gpu_runnables.h -- the reference-side binding of the B200 DVB-S receive path.
//
// A `runnable` for leansdr's own data-flow framework (reference
// src/leansdr/framework.h:124-131): the scheduler only ever calls run() and
// shutdown() (framework.h:92-95, 105-108) and the data contract is
// pipereader<T>{readable,rd,read} / pipewriter<T>{writable,wr,written}
// (framework.h:190-249).  This block replaces every runnable that
// apps/leandvb.cc:204-596 instantiates between the input pipebuf and
// p_tspackets; it owns one ldvb_handle (include/leandvb_b200.h) and moves
// whole pipebuf contents through ldvb_push()/ldvb_pull().
//
// Include AFTER "leansdr/framework.h" and "leansdr/dvb.h" of the reference tree
// (this header uses leansdr::runnable, pipebuf, pipereader, pipewriter, fail and
// tspacket from them; it defines nothing that the reference already defines).
#ifndef LEANSDR_B200_GPU_RUNNABLES_H
#define LEANSDR_B200_GPU_RUNNABLES_H

#include <stdio.h>
#include <string.h>
#include "leandvb_b200.h"
#include "leandvb_b200_tx.h"

// --const STRING as leandvb / leandvbtx spell it (leandvb.cc:1089-1112, leandvbtx.cc:256-277).
static inline int ldvb_cstln_from_name(const char *v) {
  static const char *names[] = { "BPSK", "QPSK", "8PSK", "16APSK", "32APSK", "64APSKe",
				 "16QAM", "64QAM", "256QAM" };
  for ( int k=0; k<9; ++k ) if ( !strcmp(v, names[k]) ) return k;   // = LDVB_CSTLN_*
  return -1;
}

namespace leansdr {

// Tin: complex<u8> / complex<s8> / complex<u16> / complex<s16> / complex<f32>,
// matching cfg.input_format.  Tpacket: dvb.h's tspacket (188 bytes).
template<typename Tin, typename Tpacket>
struct gpu_dvbs_receiver : runnable {
  // run() hands everything readable to the GPU at once.  It must not hold data
  // back: a step in which no pipe counter moves ends scheduler::run()
  // (framework.h:96-104), so "wait for a bigger batch" is not expressible.  Batch
  // size is therefore set by the input pipebuf (--inbuf / --buf-factor): a
  // file_reader fills it completely on every step (generic.h:51).
  gpu_dvbs_receiver(scheduler *sch, pipebuf<Tin> &_in, pipebuf<Tpacket> &_out,
		    const ldvb_config &cfg,
		    pipebuf<float> *_freq_out=NULL, pipebuf<float> *_ss_out=NULL,
		    pipebuf<float> *_mer_out=NULL, pipebuf<int> *_lock_out=NULL,
		    pipebuf<float> *_vber_out=NULL,
		    pipebuf<float> *_cnr_out=NULL,             // p_cnr, leandvb.cc:322
		    pipebuf<float[1024]> *_spectrum_out=NULL)   // p_spectrum, leandvb.cc:333-338
    : runnable(sch, "gpu_dvbs_receiver"),
      in(_in), out(_out), handle(NULL),
      last_lock(-1), rs_bits(0), rs_errs(0) {
    if ( sizeof(Tpacket) != 188 ) fail("gpu_dvbs_receiver: Tpacket must be 188 bytes");
    freq_out = opt_writer(_freq_out);
    ss_out = opt_writer(_ss_out);
    mer_out = opt_writer(_mer_out);
    lock_out = opt_writer(_lock_out);
    vber_out = opt_writer(_vber_out);
    cnr_out = opt_writer(_cnr_out);
    spectrum_out = _spectrum_out ? new pipewriter<float[1024]>(*_spectrum_out) : NULL;
    // async_push: run() returns as soon as the samples have left the pipebuf, the chain works on the
    // previous batch while file_reader refills the pipe; packets are drained on later steps.
    ldvb_config c2 = cfg;
    c2.async_push = 1;
    int rc = ldvb_create(&c2, &handle);
    if ( rc ) { fprintf(stderr, "ldvb_create: %s\n", ldvb_strerror(rc)); fail("gpu_dvbs_receiver"); }
    max_batch = cfg.max_batch;
    // The pipebufs are allocated once with new T[] (framework.h:139-141): page-lock them so that
    // the copies of ldvb_push are DMA transfers straight from in.rd().  Failure is not fatal.
    if ( ldvb_host_register(_in.buf, (size_t)(_in.end-_in.buf)*sizeof(Tin)) )
      fprintf(stderr, "gpu_dvbs_receiver: input pipebuf stays pageable (slower copies)\n");
    else registered = _in.buf;
  }

  void run() {
    // 1. Drain finished packets first (never blocks, no progress if none).
    drain();
    // 2. Feed.  No progress when starved, progress whenever possible
    //    (framework.h:96-113 detects the fixpoint on the pipe counters).
    unsigned long n = in.readable();
    if ( n > max_batch ) n = max_batch;
    if ( !n ) {
      // Starved (end of input): wait for the work in flight so that this step still makes progress
      // on the output pipe; a step in which nothing moves ends scheduler::run() (framework.h:96-104).
      // (no telemetry here: writing p_freq / p_ss / p_mer on a starved step would move pipe counters for ever)
      if ( ldvb_flush(handle) ) fail("ldvb_flush");
      drain();
      return;
    }
    int rc = ldvb_push(handle, in.rd(), n);
    if ( rc ) { fprintf(stderr, "ldvb_push: %s (%s)\n", ldvb_strerror(rc), ldvb_last_error(handle)); fail("gpu_dvbs_receiver"); }
    in.read(n);
    telemetry();
    drain();
  }

  void shutdown() {
    if ( registered ) { ldvb_host_unregister(registered); registered = NULL; }
    if ( handle ) { ldvb_destroy(handle); handle = NULL; }
  }

private:
  void *registered = NULL;
  void drain() {
    while ( 1 ) {
      unsigned long w = out.writable();
      if ( !w ) return;
      size_t got = 0;
      int rc = ldvb_pull(handle, (uint8_t*)out.wr(), w, &got);
      if ( rc ) fail("ldvb_pull");
      if ( !got ) return;
      out.written(got);
    }
  }
  // p_freq/p_ss/p_mer/p_lock/p_vber of leandvb.cc:600-616, once per batch.
  void telemetry() {
    ldvb_meas m;
    if ( ldvb_get_meas(handle, &m) ) return;
    size_t k = 0;
    float v;
    if ( freq_out && freq_out->writable() ) freq_out->write(m.freq_tap);
    if ( ss_out && ss_out->writable() ) ss_out->write(m.ss);
    if ( mer_out && mer_out->writable() ) mer_out->write(m.mer);
    if ( lock_out && m.lock != last_lock && lock_out->writable() ) { lock_out->write(m.lock); last_lock = m.lock; }
    // rate_estimator (generic.h:272-305, leandvb.cc:583-587): needs cfg.vber
    while ( vber_out && vber_out->writable() && !ldvb_pull_vber(handle, &v, 1, &k) && k ) vber_out->write(v);
    // cnr_fft / spectrum (sdr.h:1273-1404): one value / one row per second of signal
    while ( cnr_out && cnr_out->writable() && !ldvb_pull_cnr(handle, &v, 1, &k) && k ) cnr_out->write(v);
    while ( spectrum_out && spectrum_out->writable() &&
	    !ldvb_pull_spectrum(handle, (float*)spectrum_out->wr(), 1, &k) && k ) spectrum_out->written(1);
  }

  pipereader<Tin> in;
  pipewriter<Tpacket> out;
  ldvb_handle *handle;
  unsigned long max_batch;
  pipewriter<float> *freq_out, *ss_out, *mer_out, *vber_out, *cnr_out;
  pipewriter<float[1024]> *spectrum_out;
  pipewriter<int> *lock_out;
  int last_lock;
  uint64_t rs_bits, rs_errs;
};

// The transmit side: replaces every runnable that apps/leandvbtx.cc:79-197 instantiates
// between p_tspackets and the output file_writer (randomizer, rs_encoder, interleaver,
// dvb_convol, cstln_transmitter, fir_resampler, decimator, simple_agc) with one
// ldvbtx_handle (include/leandvb_b200_tx.h).  Tpacket: dvb.h's tspacket; Tout: complex<f32>.
template<typename Tpacket, typename Tout>
struct gpu_dvbs_transmitter : runnable {
  gpu_dvbs_transmitter(scheduler *sch, pipebuf<Tpacket> &_in, pipebuf<Tout> &_out,
		       const ldvbtx_config &cfg)
    : runnable(sch, "gpu_dvbs_transmitter"),
      in(_in), out(_out), handle(NULL), max_packets(cfg.max_packets) {
    if ( sizeof(Tpacket) != 188 || sizeof(Tout) != 8 ) fail("gpu_dvbs_transmitter: tspacket in, cf32 out");
    int rc = ldvbtx_create(&cfg, &handle);
    if ( rc ) { fprintf(stderr, "ldvbtx_create: %s\n", ldvb_strerror(rc)); fail("gpu_dvbs_transmitter"); }
  }

  void run() {
    // Take as many packets as the output pipe can certainly hold the samples of; no
    // progress when starved or when the output is full (framework.h:96-113).
    unsigned long n = in.readable();
    if ( n > max_packets ) n = max_packets;
    while ( n && ldvbtx_max_samples(handle, n) > out.writable() ) n /= 2;
    if ( !n ) return;
    size_t got = 0;
    int rc = ldvbtx_push(handle, (const uint8_t*)in.rd(), n, (float*)out.wr(), out.writable(), &got);
    if ( rc ) { fprintf(stderr, "ldvbtx_push: %s (%s)\n", ldvb_strerror(rc), ldvbtx_last_error(handle)); fail("gpu_dvbs_transmitter"); }
    in.read(n);
    out.written(got);
  }

  void shutdown() {
    if ( handle ) { ldvbtx_destroy(handle); handle = NULL; }
  }

private:
  pipereader<Tpacket> in;
  pipewriter<Tout> out;
  ldvbtx_handle *handle;
  unsigned long max_packets;
};

}  // namespace

#endif  // LEANSDR_B200_GPU_RUNNABLES_H
