// oracle/ref_tap.cc -- TEST INFRASTRUCTURE, not product code.
//
// Grey-box oracle: instantiates the UNMODIFIED reference runnables from
// /root/reference/src/leansdr/*.h (included, never copied) in the same order
// as the reference front end wires them (reference src/apps/leandvb.cc:204-596)
// and attaches an extra reader to every intermediate pipebuf so each stage's
// stream is written to <tapdir>/<name>.bin.  SURVEY.md section 4 verified that
// extra readers do not perturb the TS output (pipebuf supports 8 readers,
// framework.h:47,148-152).
//
// Built only in this container (oracle/Makefile -> oracle/_ref/ref_tap); the
// binary travels to the GPU box, the reference sources do not.
//
// Usage: ref_tap [leandvb-style flags] --tap-dir DIR  < IQ  > TS
//
// Dumps:
//   pp.cf32       preprocessed baseband entering cstln_receiver
//   symbols.bin   softsymbol as 4 bytes {int16 cost LE, u8 symbol, 0}
//   bytes.u8      deconvolved (or Viterbi) bytes
//   mpegbytes.u8  bit/packet aligned bytes
//   rspackets.u8  deinterleaved 204-byte packets
//   rtspackets.u8 RS-decoded, still randomised 188-byte packets
//   lock.i32, locktime.u64, freq.f32, ss.f32, mer.f32, sampled.cf32,
//   vbits.i32, verrs.i32, vber.f32, cnr.f32, spectrum.f32
//   state.txt     final private state of the stateful runnables

#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>
#include <string.h>
#include <math.h>
#include <fcntl.h>
#include <errno.h>
#include <string>

// The tap tool needs the runnables' private carry state (mu, phase, ...).
#define private public
#include "leansdr/framework.h"
#include "leansdr/generic.h"
#include "leansdr/dsp.h"
#include "leansdr/sdr.h"
#include "leansdr/dvb.h"
#include "leansdr/rs.h"
#include "leansdr/filtergen.h"
#undef private

using namespace leansdr;

template<typename T>
struct tapper : runnable {
  pipereader<T> in;
  FILE *f;
  tapper(scheduler *sch, pipebuf<T> &p, const std::string &path)
    : runnable(sch, "tap"), in(p) {
    f = fopen(path.c_str(), "wb");
    if ( !f ) fatal(path.c_str());
  }
  void run() {
    unsigned long n = in.readable();
    if ( !n ) return;
    emit(in.rd(), n);
    in.read(n);
  }
  void emit(T *p, unsigned long n) { fwrite(p, sizeof(T), n, f); }
  void shutdown() { fflush(f); }
};

// softsymbol has an uninitialised pad byte: write fields explicitly.
template<>
void tapper<softsymbol>::emit(softsymbol *p, unsigned long n) {
  for ( unsigned long i=0; i<n; ++i ) {
    unsigned char rec[4];
    rec[0] = (unsigned char)(p[i].cost & 0xff);
    rec[1] = (unsigned char)((p[i].cost >> 8) & 0xff);
    rec[2] = p[i].symbol;
    rec[3] = 0;
    fwrite(rec, 1, 4, f);
  }
}

struct opts {
  int fmt;  // 0=u8 1=s8 2=u16 3=s16 4=f32
  float float_scale, Fs, Fm, Fderot, Ftune, resample_rej, rrc_rej, rolloff, Finfo;
  int anf, decim, buf_factor, rrc_steps, sampler;
  bool cnr, resample, viterbi, hard_metric, fastlock, drift;
  code_rate fec;
  cstln_lut<256>::predef cstln;
  std::string tapdir;
  opts() : fmt(0), float_scale(1), Fs(2.4e6), Fm(2e6), Fderot(0), Ftune(0),
	   resample_rej(10), rrc_rej(10), rolloff(0.35), Finfo(5),
	   anf(1), decim(0), buf_factor(4), rrc_steps(0), sampler(1),
	   cnr(false), resample(false), viterbi(false), hard_metric(false),
	   fastlock(false), drift(false), fec(FEC12),
	   cstln(cstln_lut<256>::QPSK), tapdir("") { }
};

static int idecim(float a, float b) { int d = a/b; return d<1 ? 1 : d; }

int main(int argc, const char *argv[]) {
  opts o;
  for ( int i=1; i<argc; ++i ) {
    std::string a = argv[i];
    bool more = i+1 < argc;
    if ( a=="--u8" ) o.fmt=0; else if ( a=="--s8" ) o.fmt=1;
    else if ( a=="--u16" ) o.fmt=2; else if ( a=="--s16" ) o.fmt=3;
    else if ( a=="--f32" ) o.fmt=4;
    else if ( a=="--float-scale" && more ) o.float_scale = atof(argv[++i]);
    else if ( a=="-f" && more ) o.Fs = atof(argv[++i]);
    else if ( a=="--sr" && more ) o.Fm = atof(argv[++i]);
    else if ( a=="--anf" && more ) o.anf = atoi(argv[++i]);
    else if ( a=="--derotate" && more ) o.Fderot = atof(argv[++i]);
    else if ( a=="--tune" && more ) o.Ftune = atof(argv[++i]);
    else if ( a=="--drift" ) o.drift = true;
    else if ( a=="--cnr" ) o.cnr = true;
    else if ( a=="--resample" ) o.resample = true;
    else if ( a=="--resample-rej" && more ) o.resample_rej = atof(argv[++i]);
    else if ( a=="--decim" && more ) o.decim = atoi(argv[++i]);
    else if ( a=="--buf-factor" && more ) o.buf_factor = atoi(argv[++i]);
    else if ( a=="--viterbi" ) o.viterbi = true;
    else if ( a=="--hard-metric" ) o.hard_metric = true;
    else if ( a=="--fastlock" ) o.fastlock = true;
    else if ( a=="--rrc-steps" && more ) o.rrc_steps = atoi(argv[++i]);
    else if ( a=="--rrc-rej" && more ) o.rrc_rej = atof(argv[++i]);
    else if ( a=="--roll-off" && more ) o.rolloff = atof(argv[++i]);
    else if ( a=="--sampler" && more ) {
      std::string s = argv[++i];
      o.sampler = (s=="nearest") ? 0 : (s=="rrc") ? 2 : 1;
    }
    else if ( a=="--cr" && more ) {
      std::string s = argv[++i];
      o.fec = (s=="1/2")?FEC12:(s=="2/3")?FEC23:(s=="3/4")?FEC34:
	(s=="5/6")?FEC56:(s=="7/8")?FEC78:FEC12;
    }
    else if ( a=="--const" && more ) {
      std::string s = argv[++i];
      static const char *names[] = { "BPSK", "QPSK", "8PSK", "16APSK", "32APSK", "64APSKe",
				     "16QAM", "64QAM", "256QAM" };   // order of cstln_lut<256>::predef
      o.cstln = cstln_lut<256>::QPSK;
      for ( int k=0; k<9; ++k )
	if ( s == names[k] ) o.cstln = (cstln_lut<256>::predef)k;
    }
    else if ( a=="--tap-dir" && more ) o.tapdir = argv[++i];
    else { fprintf(stderr, "ref_tap: bad option %s\n", argv[i]); return 2; }
  }
  if ( o.tapdir.empty() ) { fprintf(stderr, "ref_tap: --tap-dir required\n"); return 2; }
  std::string T = o.tapdir + "/";

  scheduler sch;
  unsigned long BB = 4096*o.buf_factor, SY = 1024*o.buf_factor,
    BY = 2048*o.buf_factor, MB = 2448*o.buf_factor, PK = o.buf_factor,
    SL = o.buf_factor;

  // Input conversion (leandvb.cc:206-260)
  pipebuf<cf32> p_rawiq(&sch, "rawiq", BB);
  switch ( o.fmt ) {
  case 0: {
    pipebuf<cu8> *p = new pipebuf<cu8>(&sch, "stdin", BB);
    new file_reader<cu8>(&sch, 0, *p);
    new cconverter<u8,128, f32,0, 1,1>(&sch, *p, p_rawiq);
    break; }
  case 1: {
    pipebuf<cs8> *p = new pipebuf<cs8>(&sch, "stdin", BB);
    new file_reader<cs8>(&sch, 0, *p);
    new cconverter<s8,0, f32,0, 1,1>(&sch, *p, p_rawiq);
    break; }
  case 2: {
    pipebuf<cu16> *p = new pipebuf<cu16>(&sch, "stdin", BB);
    new file_reader<cu16>(&sch, 0, *p);
    new cconverter<u16,32768, f32,0, 1,1>(&sch, *p, p_rawiq);
    break; }
  case 3: {
    pipebuf<cs16> *p = new pipebuf<cs16>(&sch, "stdin", BB);
    new file_reader<cs16>(&sch, 0, *p);
    new cconverter<s16,0, f32,0, 1,1>(&sch, *p, p_rawiq);
    break; }
  default: {
    pipebuf<cf32> *p = new pipebuf<cf32>(&sch, "stdin", BB);
    new file_reader<cf32>(&sch, 0, *p);
    new scaler<float,cf32,cf32>(&sch, o.float_scale, *p, p_rawiq);
    break; }
  }
  pipebuf<cf32> *pp = &p_rawiq;

  auto_notch<f32> *r_notch = NULL;
  if ( o.anf ) {  // leandvb.cc:296-306
    pipebuf<cf32> *p = new pipebuf<cf32>(&sch, "autonotched", BB);
    r_notch = new auto_notch<f32>(&sch, *pp, *p, o.anf, 0);
    pp = p;
  }
  if ( o.Fderot ) {  // leandvb.cc:310-318
    pipebuf<cf32> *p = new pipebuf<cf32>(&sch, "derotated", BB);
    new rotator<f32>(&sch, *pp, *p, -o.Fderot/o.Fs);
    pp = p;
  }
  pipebuf<f32> p_cnr(&sch, "cnr", SL);
  cnr_fft<f32> *r_cnr = NULL;
  if ( o.cnr ) {  // leandvb.cc:322-329
    r_cnr = new cnr_fft<f32>(&sch, *pp, p_cnr, o.Fm/o.Fs);
    r_cnr->decimation = idecim(o.Fs, 1);
  }
  pipebuf<f32[1024]> *p_spectrum = new pipebuf<float[1024]>(&sch, "spectrum", SL);
  {  // leandvb.cc:333-343 (always on)
    spectrum<f32> *r = new spectrum<f32>(&sch, *pp, *p_spectrum);
    r->decimation = idecim(o.Fs, 1);
    r->kavg = 0.5;
  }
  fir_filter<cf32,float> *r_resample = NULL;
  int decim = 1;
  float Fs = o.Fs;
  if ( o.resample ) {  // leandvb.cc:353-384
    if ( o.decim ) decim = o.decim;
    else { float target = o.Fm*4; decim = Fs/target; if ( decim<1 ) decim = 1; }
    float transition = (o.Fm/2) * o.rolloff;
    int order = o.resample_rej * Fs / (22*transition);
    order = ((order+1)/2) * 2;
    pipebuf<cf32> *p = new pipebuf<cf32>(&sch, "resampled", BB);
    float *coeffs;
    float Fcut = (o.Fm/2) * (1+o.rolloff/2) / Fs;
    int ncoeffs = filtergen::lowpass(order, Fcut, &coeffs);
    filtergen::normalize_dcgain(ncoeffs, coeffs, 1);
    r_resample = new fir_filter<cf32,float>(&sch, ncoeffs, coeffs, *pp, *p, decim);
    pp = p;
    Fs /= decim;
    FILE *ft = fopen((T+"fir_taps.f32").c_str(), "wb");
    fwrite(coeffs, sizeof(float), ncoeffs, ft); fclose(ft);
  }
  if ( !o.resample && o.decim>1 ) {  // leandvb.cc:389-399
    decim = o.decim;
    pipebuf<cf32> *p = new pipebuf<cf32>(&sch, "decimated", BB);
    new decimator<cf32>(&sch, decim, *pp, *p);
    pp = p;
    Fs /= decim;
  }
  new tapper<cf32>(&sch, *pp, T+"pp.cf32");

  // Receiver (leandvb.cc:427-502)
  pipebuf<softsymbol> p_symbols(&sch, "PSK soft-symbols", SY);
  pipebuf<f32> p_freq(&sch, "freq", SL), p_ss(&sch, "SS", SL), p_mer(&sch, "MER", SL);
  pipebuf<cf32> p_sampled(&sch, "PSK symbols", BB);
  sampler_interface<f32> *sampler;
  int rrc_steps = o.rrc_steps;
  if ( o.sampler == 0 ) sampler = new nearest_sampler<float>();
  else if ( o.sampler == 1 ) sampler = new linear_sampler<float>();
  else {
    float *coeffs;
    if ( rrc_steps == 0 ) rrc_steps = max(1, (int)(64*o.Fm / Fs));
    float Frrc = Fs * rrc_steps;
    float transition = (o.Fm/2) * o.rolloff;
    int order = o.rrc_rej * Frrc / (22*transition);
    int ncoeffs = filtergen::root_raised_cosine(order, o.Fm/Frrc, o.rolloff, &coeffs);
    sampler = new fir_sampler<float,float>(ncoeffs, coeffs, rrc_steps);
    FILE *ft = fopen((T+"rrc_taps.f32").c_str(), "wb");
    fwrite(coeffs, sizeof(float), ncoeffs, ft); fclose(ft);
  }
  cstln_receiver<f32> demod(&sch, sampler, *pp, p_symbols,
			    &p_freq, &p_ss, &p_mer, &p_sampled);
  demod.cstln = make_dvbs2_constellation(o.cstln, o.fec);
  if ( o.hard_metric ) demod.cstln->harden();
  demod.set_omega(Fs/o.Fm);
  if ( o.Ftune ) demod.set_freq(o.Ftune/Fs);
  if ( o.drift ) demod.set_allow_drift(true);
  if ( o.viterbi ) demod.pll_adjustment /= 6;
  demod.meas_decimation = idecim(Fs, o.Finfo);
  if ( r_resample ) {  // leandvb.cc:506-510
    r_resample->freq_tap = &demod.freq_tap;
    r_resample->tap_multiplier = 1.0 / decim;
    r_resample->freq_tol = o.Fm/(Fs*decim) * 0.1;
  }
  if ( r_cnr ) {
    r_cnr->freq_tap = &demod.freq_tap;
    r_cnr->tap_multiplier = 1.0 / decim;
  }
  new tapper<softsymbol>(&sch, p_symbols, T+"symbols.bin");
  new tapper<f32>(&sch, p_freq, T+"freq.f32");
  new tapper<f32>(&sch, p_ss, T+"ss.f32");
  new tapper<f32>(&sch, p_mer, T+"mer.f32");
  new tapper<cf32>(&sch, p_sampled, T+"sampled.cf32");
  new tapper<f32>(&sch, p_cnr, T+"cnr.f32");
  new tapper<f32[1024]>(&sch, *p_spectrum, T+"spectrum.f32");

  // Deconvolution and sync (leandvb.cc:527-566)
  pipebuf<u8> p_bytes(&sch, "bytes", BY);
  deconvol_sync_simple *r_deconv = NULL;
  code_rate fec = o.fec;
  if ( o.viterbi ) {
    if ( fec==FEC23 && (demod.cstln->nsymbols==4 || demod.cstln->nsymbols==64) ) fec = FEC46;
    viterbi_sync *r = new viterbi_sync(&sch, p_symbols, p_bytes, demod.cstln, fec);
    if ( o.fastlock ) r->resync_period = 1;
  } else {
    r_deconv = make_deconvol_sync_simple(&sch, p_symbols, p_bytes, fec);
    r_deconv->fastlock = o.fastlock;
  }
  new tapper<u8>(&sch, p_bytes, T+"bytes.u8");
  pipebuf<u8> p_mpegbytes(&sch, "mpegbytes", MB);
  pipebuf<int> p_lock(&sch, "lock", SL);
  pipebuf<u32> p_locktime(&sch, "locktime", PK);
  mpeg_sync<u8,0> *r_sync = new mpeg_sync<u8,0>(&sch, p_bytes, p_mpegbytes, r_deconv,
						&p_lock, &p_locktime);
  r_sync->fastlock = o.fastlock;
  new tapper<u8>(&sch, p_mpegbytes, T+"mpegbytes.u8");
  new tapper<int>(&sch, p_lock, T+"lock.i32");
  new tapper<u32>(&sch, p_locktime, T+"locktime.u64");

  pipebuf< rspacket<u8> > p_rspackets(&sch, "RS-enc packets", PK);
  deinterleaver<u8> r_deinter(&sch, p_mpegbytes, p_rspackets);
  new tapper< rspacket<u8> >(&sch, p_rspackets, T+"rspackets.u8");

  pipebuf<int> p_vbitcount(&sch, "Bits processed", PK);
  pipebuf<int> p_verrcount(&sch, "Bits corrected", PK);
  pipebuf<tspacket> p_rtspackets(&sch, "rand TS packets", PK);
  rs_decoder<u8,0> r_rsdec(&sch, p_rspackets, p_rtspackets, &p_vbitcount, &p_verrcount);
  new tapper<tspacket>(&sch, p_rtspackets, T+"rtspackets.u8");
  new tapper<int>(&sch, p_vbitcount, T+"vbits.i32");
  new tapper<int>(&sch, p_verrcount, T+"verrs.i32");

  pipebuf<float> p_vber(&sch, "VBER", SL);
  rate_estimator<float> r_vber(&sch, p_verrcount, p_vbitcount, p_vber);
  r_vber.sample_size = o.Fm/2;
  if ( r_vber.sample_size < 50000 ) r_vber.sample_size = 50000;
  new tapper<float>(&sch, p_vber, T+"vber.f32");

  pipebuf<tspacket> p_tspackets(&sch, "TS packets", PK);
  derandomizer r_derand(&sch, p_rtspackets, p_tspackets);
  file_writer<tspacket> r_stdout(&sch, p_tspackets, 1);

  sch.run();
  sch.shutdown();

  // Final carry state of the serial stages.
  FILE *fs = fopen((T+"state.txt").c_str(), "w");
  fprintf(fs, "rx.mu %a\nrx.phase %a\nrx.freqw %a\nrx.est_insp %a\nrx.agc_gain %a\n"
	  "rx.est_sp %a\nrx.est_ep %a\nrx.meas_count %lu\nrx.min_freqw %a\nrx.max_freqw %a\n",
	  demod.mu, demod.phase, demod.freqw, demod.est_insp, demod.agc_gain,
	  demod.est_sp, demod.est_ep, demod.meas_count, demod.min_freqw, demod.max_freqw);
  for ( int k=0; k<3; ++k )
    fprintf(fs, "rx.hist%d %a %a %a %a\n", k, demod.hist[k].p.re, demod.hist[k].p.im,
	    demod.hist[k].c.re, demod.hist[k].c.im);
  fprintf(fs, "rx.in_total_read %lu\nrx.out_total_written %lu\n",
	  pp->total_read, p_symbols.total_written);
  if ( r_notch ) {
    fprintf(fs, "notch.phase %d\nnotch.gain %a\n", r_notch->phase, r_notch->gain);
    for ( int s=0; s<r_notch->nslots; ++s )
      fprintf(fs, "notch.slot%d %d %a %a\n", s, r_notch->slots[s].i,
	      r_notch->slots[s].estim.re, r_notch->slots[s].estim.im);
  }
  if ( r_resample ) fprintf(fs, "fir.current_freq %a\n", r_resample->current_freq);
  if ( r_deconv ) fprintf(fs, "deconv.locked %d\ndeconv.skip %d\n",
			  (int)(r_deconv->locked - r_deconv->syncs), r_deconv->skip);
  fprintf(fs, "sync.bitphase %d\nsync.polarity %d\nsync.synchronized %d\nsync.phase8 %d\n"
	  "sync.locktime %lu\nsync.lock_timeleft %lu\n",
	  r_sync->bitphase, (int)r_sync->polarity, (int)r_sync->synchronized,
	  r_sync->phase8, r_sync->locktime, r_sync->lock_timeleft);
  fprintf(fs, "derand.pos %d\n", (int)(r_derand.pos - r_derand.pattern));
  fclose(fs);
  return 0;
}
