// oracle/ref_fastlock.cc -- TEST INFRASTRUCTURE, not product code.
//
// Drives the UNMODIFIED reference deconvol_sync (dvb.h:122-476) and mpeg_sync
// (dvb.h:712-891) with --fastlock semantics under a CONTROLLED schedule: the
// symbol stream is fed W symbols at a time, deconvol_sync::run() is called once
// per feed (what it does then depends on the window it sees, dvb.h:414-454) and
// mpeg_sync::run() until it stops moving.  oracle/dvbs_oracle.c must reproduce
// the same bytes when it is given the same windows (tests/test_oracle_cpu.py).
//
// Usage: ref_fastlock FEC W SYNC_PERIOD FASTLOCK symbols.bin out_prefix
//   FEC: 0=1/2 1=2/3 3=3/4 4=5/6 5=7/8 (code_rate order, dvb.h:36-40)
//   writes out_prefix.bytes, out_prefix.mpeg, out_prefix.state (locked, skip per run)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <string>

#define private public
#include "leansdr/framework.h"
#include "leansdr/generic.h"
#include "leansdr/dsp.h"
#include "leansdr/sdr.h"
#include "leansdr/dvb.h"
#undef private

using namespace leansdr;

int main(int argc, char **argv) {
  if ( argc != 7 ) { fprintf(stderr, "usage: ref_fastlock FEC W SYNC_PERIOD FASTLOCK symbols.bin out_prefix\n"); return 1; }
  int fec = atoi(argv[1]);
  unsigned long W = strtoul(argv[2], NULL, 0);
  int period = atoi(argv[3]);
  bool fastlock = atoi(argv[4]) != 0;
  FILE *fi = fopen(argv[5], "rb");
  if ( !fi ) fatal(argv[5]);
  std::string pre = argv[6];
  FILE *fb = fopen((pre+".bytes").c_str(), "wb"); FILE *fm = fopen((pre+".mpeg").c_str(), "wb");
  FILE *fst = fopen((pre+".state").c_str(), "w");
  scheduler sch;
  const unsigned long BIG = 1 << 22;
  pipebuf<softsymbol> p_sym(&sch, "symbols", BIG);
  pipebuf<u8> p_bytes(&sch, "bytes", BIG);
  pipebuf<u8> p_mpeg(&sch, "mpegbytes", BIG);
  pipebuf<int> p_lock(&sch, "lock", 1 << 16);
  deconvol_sync_simple *r_deconv = make_deconvol_sync_simple(&sch, p_sym, p_bytes, (code_rate)fec);
  r_deconv->fastlock = fastlock;
  mpeg_sync<u8,0> r_sync(&sch, p_bytes, p_mpeg, fastlock ? NULL : r_deconv, &p_lock, NULL);
  r_sync.fastlock = fastlock;
  r_sync.resync_period = period;
  pipewriter<softsymbol> w_sym(p_sym);
  pipereader<u8> rd_bytes(p_bytes), rd_mpeg(p_mpeg);   // extra readers: the dumps
  pipereader<int> rd_lock(p_lock);
  unsigned char buf[4];
  bool eof = false;
  while ( !eof ) {
    unsigned long fed = 0;
    while ( fed < W && w_sym.writable() ) {
      if ( fread(buf, 1, 4, fi) != 4 ) { eof = true; break; }
      softsymbol s; memset(&s, 0, sizeof s);
      s.cost = (int16_t)(buf[0] | (buf[1] << 8)); s.symbol = buf[2];
      w_sym.write(s);
      ++fed;
    }
    r_deconv->run();
    fprintf(fst, "%d %d\n", (int)(r_deconv->locked - r_deconv->syncs), r_deconv->skip);
    unsigned long nb = rd_bytes.readable();
    fwrite(rd_bytes.rd(), 1, nb, fb); rd_bytes.read(nb);
    for ( int guard = 0; guard < 1000000; ++guard ) {
      unsigned long before = p_bytes.total_read + p_mpeg.total_written;
      r_sync.run();
      unsigned long nm = rd_mpeg.readable();
      fwrite(rd_mpeg.rd(), 1, nm, fm); rd_mpeg.read(nm);
      rd_lock.read(rd_lock.readable());
      if ( p_bytes.total_read + p_mpeg.total_written == before ) break;
    }
  }
  fclose(fb); fclose(fm); fclose(fst);
  return 0;
}
