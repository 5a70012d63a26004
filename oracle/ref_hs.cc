// oracle/ref_hs.cc -- TEST INFRASTRUCTURE, not product code.
//
// The --hs receive path of the reference (apps/leandvb.cc:727-969) with an extra reader on
// every pipebuf: UNMODIFIED fast_qpsk_receiver<u8>, dvb_deconvol_sync_hard, mpeg_sync
// (fastlock), deinterleaver, rs_decoder, derandomizer from /root/reference/src/leansdr/*.h.
//
// Usage: ref_hs FS FM RESYNC_PERIOD out_prefix  < u8 IQ  > TS
//   dumps out_prefix.symbols (u8), .bytes, .mpeg
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <string>

#include "leansdr/framework.h"
#include "leansdr/generic.h"
#include "leansdr/dsp.h"
#include "leansdr/sdr.h"
#include "leansdr/dvb.h"
#include "leansdr/rs.h"

using namespace leansdr;

template<typename T>
struct dumper : runnable {
  pipereader<T> in;
  FILE *f;
  dumper(scheduler *sch, pipebuf<T> &p, const std::string &path) : runnable(sch, "dump"), in(p) {
    f = fopen(path.c_str(), "wb");
    if ( !f ) fatal(path.c_str());
  }
  void run() {
    unsigned long n = in.readable();
    if ( !n ) return;
    fwrite(in.rd(), sizeof(T), n, f);
    in.read(n);
  }
  void shutdown() { fclose(f); }
};

int main(int argc, char **argv) {
  if ( argc != 5 ) { fprintf(stderr, "usage: ref_hs FS FM RESYNC_PERIOD out_prefix\n"); return 1; }
  float Fs = atof(argv[1]), Fm = atof(argv[2]);
  int period = atoi(argv[3]);
  std::string pre = argv[4];
  scheduler sch;
  int bf = 4;
  pipebuf<cu8> p_rawiq(&sch, "rawiq", 4096*bf);
  file_reader<cu8> r_stdin(&sch, 0, p_rawiq);
  pipebuf<u8> p_symbols(&sch, "PSK hard symbols", 1024*bf);
  fast_qpsk_receiver<u8> demod(&sch, p_rawiq, p_symbols);
  demod.set_omega(Fs/Fm);
  pipebuf<u8> p_bytes(&sch, "bytes", 2048*bf);
  dvb_deconvol_sync_hard r_deconv(&sch, p_symbols, p_bytes);
  r_deconv.resync_period = period;
  pipebuf<u8> p_mpegbytes(&sch, "mpegbytes", 2448*bf);
  pipebuf<int> p_lock(&sch, "lock", bf);
  mpeg_sync<u8,0> r_sync(&sch, p_bytes, p_mpegbytes, NULL, &p_lock, NULL);
  r_sync.fastlock = true;
  r_sync.resync_period = period;
  pipebuf< rspacket<u8> > p_rspackets(&sch, "RS-enc packets", bf);
  deinterleaver<u8> r_deinter(&sch, p_mpegbytes, p_rspackets);
  pipebuf<tspacket> p_rtspackets(&sch, "rand TS packets", bf);
  rs_decoder<u8,0> r_rsdec(&sch, p_rspackets, p_rtspackets, NULL, NULL);
  pipebuf<tspacket> p_tspackets(&sch, "TS packets", bf);
  derandomizer r_derand(&sch, p_rtspackets, p_tspackets);
  file_writer<tspacket> r_stdout(&sch, p_tspackets, 1);
  dumper<u8> d1(&sch, p_symbols, pre + ".symbols"), d2(&sch, p_bytes, pre + ".bytes"), d3(&sch, p_mpegbytes, pre + ".mpeg");
  dumper<int> d4(&sch, p_lock, pre + ".lock");
  sch.run();
  sch.shutdown();
  return 0;
}
