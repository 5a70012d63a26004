// oracle/ref_tables.cc -- TEST INFRASTRUCTURE, not product code.
//
// Dumps the constant tables that the reference builds at start-up, straight
// from the unmodified reference headers, so that the oracle restatement and
// the product's host-side table builders can be checked entry by entry:
//   cstln_<name>.bin  256x256 x {int16 cost, u8 symbol, u8 0, int16 phase_error, 2x0}
//                     (reference src/leansdr/sdr.h:529-560) + symbols
//   trig16.f32        65536 x {cos, sin}       (math.h:95-111)
//   rs_exp.u8 rs_log.u8 rs_gen.u8              (rs.h:47-108)
//   deconv_<cr>.u64   deconv[] / deconv2[]     (dvb.h:205-292)
//   trellis_<cr>.bin  NSTATES x NCS x {pred, us} (viterbi.h:61-92)
//   vitmap_<cr>.u8    viterbi_sync hypothesis maps (dvb.h:1336-1351)
//   derand.u8         1504-byte PRBS pattern   (dvb.h:1116-1129)
//   lowpass_*.f32 / rrc_*.f32 filtergen outputs for the bench configs
//
// Usage: ref_tables OUTDIR

#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>
#include <string.h>
#include <math.h>
#include <string>

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

static std::string D;

static FILE *out(const std::string &name) {
  FILE *f = fopen((D+"/"+name).c_str(), "wb");
  if ( !f ) fatal(name.c_str());
  return f;
}

static void dump_cstln(const char *name, cstln_lut<256>::predef c, code_rate cr,
		       bool harden=false) {
  cstln_lut<256> *l = make_dvbs2_constellation(c, cr);
  if ( harden ) l->harden();
  FILE *f = out(std::string("cstln_")+name+".bin");
  for ( int i=0; i<256; ++i )
    for ( int q=0; q<256; ++q ) {
      cstln_lut<256>::result *r = &l->lut[i][q];
      int16_t rec[4] = { r->ss.cost, (int16_t)r->ss.symbol, r->phase_error, 0 };
      fwrite(rec, 2, 4, f);
    }
  fclose(f);
  f = out(std::string("cstln_")+name+"_symbols.s8");
  for ( int s=0; s<l->nsymbols; ++s ) {
    signed char p[2] = { l->symbols[s].re, l->symbols[s].im };
    fwrite(p, 1, 2, f);
  }
  fclose(f);
}

template<typename TR>
static void dump_trellis(const char *name, const uint16_t *polys, int nstates, int ncs) {
  TR *t = new TR();
  t->init_convolutional(polys);
  FILE *f = out(std::string("trellis_")+name+".bin");
  for ( int s=0; s<nstates; ++s )
    for ( int cs=0; cs<ncs; ++cs ) {
      // `us` of a branch that does not exist (pred == NOSTATE) is never written by the
      // reference (viterbi.h:54-58): dump 0 there so that the file is reproducible.
      unsigned char rec[2] = { t->states[s].branches[cs].pred, t->states[s].branches[cs].us };
      if ( rec[0] == TR::NOSTATE ) rec[1] = 0;
      fwrite(rec, 1, 2, f);
    }
  fclose(f);
}

static void dump_deconv(const char *name, code_rate cr) {
  scheduler sch;
  pipebuf<softsymbol> pi(&sch, "i", 4096);
  pipebuf<u8> po(&sch, "o", 4096);
  deconvol_sync_simple *d = make_deconvol_sync_simple(&sch, pi, po, cr);
  FILE *f = out(std::string("deconv_")+name+".u64");
  uint64_t hdr[2] = { (uint64_t)d->punctperiod, (uint64_t)d->punctweight };
  fwrite(hdr, 8, 2, f);
  fwrite(d->deconv, 8, d->punctperiod, f);
  fwrite(d->deconv2, 8, d->punctperiod, f);
  for ( int s=0; s<4; ++s ) {
    uint64_t l[4] = { d->syncs[s].lut[0][0], d->syncs[s].lut[0][1],
		      d->syncs[s].lut[1][0], d->syncs[s].lut[1][1] };
    fwrite(l, 8, 4, f);
  }
  fclose(f);
}

static void dump_vitmap(const char *name, cstln_lut<256>::predef c, code_rate cr) {
  scheduler sch;
  pipebuf<softsymbol> pi(&sch, "i", 4096);
  pipebuf<u8> po(&sch, "o", 4096);
  cstln_lut<256> *l = make_dvbs2_constellation(c, cr);
  viterbi_sync *v = new viterbi_sync(&sch, pi, po, l, cr);
  FILE *f = out(std::string("vitmap_")+name+".u8");
  unsigned char hdr[4] = { (unsigned char)v->nsyncs, (unsigned char)v->nshifts,
			   (unsigned char)v->bits_per_symbol, (unsigned char)l->nsymbols };
  fwrite(hdr, 1, 4, f);
  for ( int s=0; s<v->nsyncs; ++s ) {
    unsigned char sh = v->syncs[s].shift;
    fwrite(&sh, 1, 1, f);
    fwrite(v->syncs[s].map, 1, l->nsymbols, f);
  }
  fclose(f);
}

static void dump_lowpass(const char *name, float Fs, float Fm, float rolloff, float rej) {
  float transition = (Fm/2) * rolloff;
  int order = rej * Fs / (22*transition);
  order = ((order+1)/2) * 2;
  float *coeffs;
  float Fcut = (Fm/2) * (1+rolloff/2) / Fs;
  int n = filtergen::lowpass(order, Fcut, &coeffs);
  filtergen::normalize_dcgain(n, coeffs, 1);
  FILE *f = out(std::string("lowpass_")+name+".f32");
  fwrite(coeffs, 4, n, f);
  fclose(f);
}

static void dump_rrc(const char *name, float Fs, float Fm, float rolloff, float rej) {
  int steps = max(1, (int)(64*Fm / Fs));
  float Frrc = Fs * steps;
  float transition = (Fm/2) * rolloff;
  int order = rej * Frrc / (22*transition);
  float *coeffs;
  int n = filtergen::root_raised_cosine(order, Fm/Frrc, rolloff, &coeffs);
  FILE *f = out(std::string("rrc_")+name+".f32");
  fwrite(coeffs, 4, n, f);
  fclose(f);
}

int main(int argc, char **argv) {
  if ( argc != 2 ) { fprintf(stderr, "usage: ref_tables OUTDIR\n"); return 2; }
  D = argv[1];

  dump_cstln("qpsk", cstln_lut<256>::QPSK, FEC12);
  dump_cstln("qpsk_hard", cstln_lut<256>::QPSK, FEC12, true);
  dump_cstln("bpsk", cstln_lut<256>::BPSK, FEC12);
  dump_cstln("8psk", cstln_lut<256>::PSK8, FEC12);
  dump_cstln("16apsk23", cstln_lut<256>::APSK16, FEC23);
  dump_cstln("16apsk34", cstln_lut<256>::APSK16, FEC34);
  dump_cstln("16apsk56", cstln_lut<256>::APSK16, FEC56);
  dump_cstln("32apsk34", cstln_lut<256>::APSK32, FEC34);
  dump_cstln("32apsk56", cstln_lut<256>::APSK32, FEC56);
  dump_cstln("64apske", cstln_lut<256>::APSK64E, FEC34);
  dump_cstln("16qam", cstln_lut<256>::QAM16, FEC12);
  dump_cstln("64qam", cstln_lut<256>::QAM64, FEC12);
  dump_cstln("256qam", cstln_lut<256>::QAM256, FEC12);
  dump_cstln("16apsk34_hard", cstln_lut<256>::APSK16, FEC34, true);

  { trig16 *t = new trig16();
    FILE *f = out("trig16.f32"); fwrite(t->lut, 8, 65536, f); fclose(f); }

  { rs_engine *rs = new rs_engine();
    // lut_log[0] is never written by the reference (rs.h:53-60); pin it to 0.
    FILE *f = out("rs_exp.u8"); fwrite(rs->gf.lut_exp, 1, 511, f); fclose(f);
    f = out("rs_log.u8"); unsigned char z = 0; fwrite(&z, 1, 1, f);
    fwrite(rs->gf.lut_log+1, 1, 255, f); fclose(f);
    f = out("rs_gen.u8"); fwrite(rs->G, 1, 17, f); fclose(f); }

  dump_deconv("12", FEC12); dump_deconv("23", FEC23); dump_deconv("34", FEC34);
  dump_deconv("56", FEC56); dump_deconv("78", FEC78);

  dump_trellis<viterbi_sync::trellis_12>("12", polys_fec12, 64, 4);
  dump_trellis<viterbi_sync::trellis_46>("46", polys_fec46, 64, 64);
  dump_trellis<viterbi_sync::trellis_34>("34", polys_fec34, 64, 16);
  dump_trellis<viterbi_sync::trellis_56>("56", polys_fec56, 64, 64);
  dump_trellis<viterbi_sync::trellis_78>("78", polys_fec78, 64, 256);

  dump_vitmap("qpsk12", cstln_lut<256>::QPSK, FEC12);
  dump_vitmap("qpsk34", cstln_lut<256>::QPSK, FEC34);
  dump_vitmap("qpsk78", cstln_lut<256>::QPSK, FEC78);
  dump_vitmap("8psk23", cstln_lut<256>::PSK8, FEC23);
  dump_vitmap("16apsk34", cstln_lut<256>::APSK16, FEC34);
  dump_vitmap("16qam34", cstln_lut<256>::QAM16, FEC34);
  dump_vitmap("64qam46", cstln_lut<256>::QAM64, FEC46);
  dump_vitmap("256qam78", cstln_lut<256>::QAM256, FEC78);

  { scheduler sch;
    pipebuf<tspacket> a(&sch, "a", 4), b(&sch, "b", 4);
    derandomizer *d = new derandomizer(&sch, a, b);
    FILE *f = out("derand.u8"); fwrite(d->pattern, 1, 188*8, f); fclose(f); }

  { // fast_qpsk_receiver<u8>::init_lookup_tables (sdr.h:1144-1164), packed like the B200 library uploads them:
    // polar = angle | radius << 16 (u32), rect / sincos = re | im << 8 (u16)
    scheduler sch;
    pipebuf<cu8> pi(&sch, "i", 4096);
    pipebuf<u8> po(&sch, "o", 4096);
    fast_qpsk_receiver<u8> *r = new fast_qpsk_receiver<u8>(&sch, pi, po);
    FILE *f = out("hs_polar.u32");
    for ( int i=0; i<256; ++i )
      for ( int q=0; q<256; ++q ) {
	uint32_t w = (uint32_t)r->lut_polar[i][q].a | ((uint32_t)r->lut_polar[i][q].r << 16);
	fwrite(&w, 4, 1, f);
      }
    fclose(f);
    f = out("hs_rect.u16");
    for ( int a=0; a<256; ++a )
      for ( int k=0; k<256; ++k ) {
	uint16_t w = (uint16_t)(r->lut_rect[a][k].re | (r->lut_rect[a][k].im << 8));
	fwrite(&w, 2, 1, f);
      }
    fclose(f);
    f = out("hs_sincos.u16");
    for ( int a=0; a<65536; ++a ) {
      uint16_t w = (uint16_t)(r->lut_sincos[a].re | (r->lut_sincos[a].im << 8));
      fwrite(&w, 2, 1, f);
    }
    fclose(f); }

  dump_lowpass("fs2.4_sr2", 2.4e6, 2e6, 0.35, 10);
  dump_lowpass("fs9.6_sr2", 9.6e6, 2e6, 0.35, 10);
  dump_lowpass("fs240_sr2", 240e6, 2e6, 0.35, 10);
  dump_rrc("fs2.4_sr2", 2.4e6, 2e6, 0.35, 10);
  dump_rrc("fs4_sr2", 4e6, 2e6, 0.35, 10);
  dump_lowpass("fs9.6_sr2_rej20", 9.6e6, 2e6, 0.35, 20);
  dump_lowpass("fs2.4_sr2_ro02", 2.4e6, 2e6, 0.2, 10);
  dump_lowpass("fs55_sr27.5", 55e6, 27.5e6, 0.35, 10);
  dump_rrc("fs2.4_sr2_ro02", 2.4e6, 2e6, 0.2, 10);
  dump_rrc("fs8_sr2_rej5", 8e6, 2e6, 0.35, 5);
  return 0;
}
