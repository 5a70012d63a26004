// leandvb_gpu.cc -- `leandvb` with the receive chain on a B200.
//
// Same command line as the reference front end for the options that reach the
// hot path (apps/leandvb.cc:1064-1215); the graph is
//   file_reader<T>(stdin) -> gpu_dvbs_receiver -> file_writer<tspacket>(stdout)
// on the reference's unmodified scheduler / pipebuf / file_reader / file_writer
// (framework.h, generic.h).  Built against the reference tree:
//   g++ -O2 -I$REF/src -I$REPO/include -I$REPO/leansdr_b200/host leandvb_gpu.cc \
//       -L$REPO/leansdr_b200 -lleandvb_b200 -o leandvb_gpu
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>
#include <string.h>
#include <math.h>
#include <fcntl.h>
#include <time.h>

#include "leansdr/framework.h"
#include "leansdr/generic.h"
#include "leansdr/dsp.h"
#include "leansdr/sdr.h"
#include "leansdr/dvb.h"
#include "gpu_runnables.h"

using namespace leansdr;

template<typename Tin>
static int run_chain(const ldvb_config &cfg, unsigned long inbuf, bool info, bool timing) {
  scheduler sch;
  pipebuf<Tin> p_stdin(&sch, "stdin", inbuf);
  pipebuf<tspacket> p_ts(&sch, "TS packets", 1<<16);
  pipebuf<float> p_freq(&sch, "freq", 64), p_ss(&sch, "SS", 64), p_mer(&sch, "MER", 64), p_vber(&sch, "VBER", 64);
  pipebuf<int> p_lock(&sch, "lock", 64);
  file_reader<Tin> r_stdin(&sch, 0, p_stdin);
  gpu_dvbs_receiver<Tin, tspacket> r_gpu(&sch, p_stdin, p_ts, cfg, &p_freq, &p_ss, &p_mer, &p_lock, &p_vber);
  file_writer<tspacket> r_stdout(&sch, p_ts, 1);
  if ( info ) {
    file_printer<float> *pf = new file_printer<float>(&sch, "FREQ %.0f\n", p_freq, 2);
    pf->scale = cfg.Fs;
    new file_printer<float>(&sch, "SS %f\n", p_ss, 2);
    new file_printer<float>(&sch, "MER %.1f\n", p_mer, 2);
    new file_printer<int>(&sch, "LOCK %d\n", p_lock, 2);
    new file_printer<float>(&sch, "VBER %.6f\n", p_vber, 2);
  }
  // --gpu-timing: wall clock of the scheduler loop alone (CUDA start-up and ldvb_create lie in front of it)
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  sch.run();
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if ( timing )
    fprintf(stderr, "LDVB_TIMING samples=%lu seconds=%.6f\n", (unsigned long)p_stdin.total_written,
	    (t1.tv_sec-t0.tv_sec) + 1e-9*(t1.tv_nsec-t0.tv_nsec));
  sch.shutdown();
  return 0;
}

int main(int argc, const char *argv[]) {
  ldvb_config cfg;
  ldvb_config_default(&cfg);
  cfg.max_batch = 1<<24;
  cfg.rx_mode = LDVB_RX_FAST;
  bool info = false, timing = false;
  for ( int i=1; i<argc; ++i ) {
    const char *a = argv[i];
    bool more = i+1 < argc;
    if      ( !strcmp(a,"--u8") ) cfg.input_format = LDVB_FMT_U8;
    else if ( !strcmp(a,"--s8") ) cfg.input_format = LDVB_FMT_S8;
    else if ( !strcmp(a,"--u16") ) cfg.input_format = LDVB_FMT_U16;
    else if ( !strcmp(a,"--s16") ) cfg.input_format = LDVB_FMT_S16;
    else if ( !strcmp(a,"--f32") ) cfg.input_format = LDVB_FMT_F32;
    else if ( !strcmp(a,"--float-scale") && more ) cfg.float_scale = atof(argv[++i]);
    else if ( !strcmp(a,"-f") && more ) cfg.Fs = atof(argv[++i]);
    else if ( !strcmp(a,"--sr") && more ) cfg.Fm = atof(argv[++i]);
    else if ( !strcmp(a,"--anf") && more ) cfg.anf = atoi(argv[++i]);
    else if ( !strcmp(a,"--derotate") && more ) cfg.Fderot = atof(argv[++i]);
    else if ( !strcmp(a,"--resample") ) cfg.resample = 1;
    else if ( !strcmp(a,"--resample-rej") && more ) cfg.resample_rej = atof(argv[++i]);
    else if ( !strcmp(a,"--decim") && more ) cfg.decim = atoi(argv[++i]);
    else if ( !strcmp(a,"--tune") && more ) cfg.Ftune = atof(argv[++i]);
    else if ( !strcmp(a,"--drift") ) cfg.allow_drift = 1;
    else if ( !strcmp(a,"--roll-off") && more ) cfg.rolloff = atof(argv[++i]);
    else if ( !strcmp(a,"--hard-metric") ) cfg.hard_metric = 1;
    else if ( !strcmp(a,"--fastlock") ) cfg.fastlock = 1;
    else if ( !strcmp(a,"--hs") ) cfg.hs = 1;
    else if ( !strcmp(a,"--viterbi") ) cfg.viterbi = 1;
    else if ( !strcmp(a,"--standard") && more ) ++i;       // DVB-S only
    else if ( !strcmp(a,"--const") && more ) cfg.constellation = ldvb_cstln_from_name(argv[++i]);
    else if ( !strcmp(a,"--cr") && more ) {
      const char *v = argv[++i];
      cfg.fec = !strcmp(v,"1/2") ? LDVB_FEC12 : !strcmp(v,"2/3") ? LDVB_FEC23 : !strcmp(v,"3/4") ? LDVB_FEC34 :
	!strcmp(v,"5/6") ? LDVB_FEC56 : !strcmp(v,"7/8") ? LDVB_FEC78 : -1;
    }
    else if ( !strcmp(a,"--sampler") && more ) {
      const char *v = argv[++i];
      cfg.sampler = !strcmp(v,"nearest") ? LDVB_SAMP_NEAREST : !strcmp(v,"rrc") ? LDVB_SAMP_RRC : LDVB_SAMP_LINEAR;
    }
    else if ( !strcmp(a,"--gpu-exact") ) cfg.rx_mode = LDVB_RX_EXACT;
    else if ( !strcmp(a,"--gpu-batch") && more ) cfg.max_batch = strtoull(argv[++i], NULL, 0);
    else if ( !strcmp(a,"--gpu-device") && more ) cfg.device = atoi(argv[++i]);
    else if ( !strcmp(a,"--gpu-timing") ) timing = true;
    else if ( !strcmp(a,"--fd-info") && more ) { info = (atoi(argv[++i]) == 2); cfg.vber = 1; }
    else { fprintf(stderr, "leandvb_gpu: unsupported option %s\n", a); return 1; }
  }
  unsigned long inbuf = cfg.max_batch;
  switch ( cfg.input_format ) {
  case LDVB_FMT_U8:  return run_chain< complex<u8> >(cfg, inbuf, info, timing);
  case LDVB_FMT_S8:  return run_chain< complex<s8> >(cfg, inbuf, info, timing);
  case LDVB_FMT_U16: return run_chain< complex<u16> >(cfg, inbuf, info, timing);
  case LDVB_FMT_S16: return run_chain< complex<s16> >(cfg, inbuf, info, timing);
  default:           return run_chain< complex<f32> >(cfg, inbuf, info, timing);
  }
}
