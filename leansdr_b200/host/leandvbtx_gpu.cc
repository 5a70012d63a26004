// leandvbtx_gpu.cc -- `leandvbtx` with the transmit chain on a B200.
//
// Same command line as the reference front end (apps/leandvbtx.cc:258-297); the graph is
//   file_reader<tspacket>(stdin) -> gpu_dvbs_transmitter -> file_writer<cf32>(stdout)
// on the reference's unmodified scheduler / pipebuf / file_reader / file_writer.
// Built against the reference tree (oracle/Makefile, target _ref/leandvbtx_gpu).
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>
#include <string.h>
#include <math.h>

#include "leansdr/framework.h"
#include "leansdr/generic.h"
#include "leansdr/dsp.h"
#include "leansdr/sdr.h"
#include "leansdr/dvb.h"
#include "gpu_runnables.h"

using namespace leansdr;

int main(int argc, const char *argv[]) {
  ldvbtx_config cfg;
  ldvbtx_config_default(&cfg);
  cfg.max_packets = 1<<14;
  for ( int i=1; i<argc; ++i ) {
    const char *a = argv[i];
    bool more = i+1 < argc;
    if ( !strcmp(a,"--const") && more ) {
      cfg.constellation = ldvb_cstln_from_name(argv[++i]);
    }
    else if ( !strcmp(a,"--cr") && more ) {
      const char *v = argv[++i];
      cfg.fec = !strcmp(v,"1/2") ? LDVB_FEC12 : !strcmp(v,"2/3") ? LDVB_FEC23 : !strcmp(v,"3/4") ? LDVB_FEC34 :
	!strcmp(v,"5/6") ? LDVB_FEC56 : !strcmp(v,"7/8") ? LDVB_FEC78 : -1;
    }
    else if ( !strcmp(a,"-f") && more ) {               // leandvbtx.cc:279-283
      ++i;
      cfg.decim = 1;
      if ( sscanf(argv[i], "%d/%d", &cfg.interp, &cfg.decim) < 1 ) { fprintf(stderr, "bad -f\n"); return 1; }
    }
    else if ( !strcmp(a,"--roll-off") && more ) cfg.rolloff = atof(argv[++i]);
    else if ( !strcmp(a,"--rrc-rej") && more ) cfg.rrc_rej = atof(argv[++i]);
    else if ( !strcmp(a,"--power") && more ) { strncpy(cfg.power_db, argv[++i], sizeof cfg.power_db - 1); }
    else if ( !strcmp(a,"--agc") ) cfg.agc = 1;
    else if ( !strcmp(a,"--f32") ) ;
    else if ( !strcmp(a,"--gpu-device") && more ) cfg.device = atoi(argv[++i]);
    else if ( !strcmp(a,"--gpu-batch") && more ) cfg.max_packets = strtoull(argv[++i], NULL, 0);
    else { fprintf(stderr, "leandvbtx_gpu: unsupported option %s\n", a); return 1; }
  }
  scheduler sch;
  pipebuf<tspacket> p_ts(&sch, "TS packets", cfg.max_packets);
  pipebuf<cf32> p_iq(&sch, "IQ", 1<<26);
  file_reader<tspacket> r_stdin(&sch, 0, p_ts);
  gpu_dvbs_transmitter<tspacket, cf32> r_gpu(&sch, p_ts, p_iq, cfg);
  file_writer<cf32> r_stdout(&sch, p_iq, 1);
  sch.run();
  sch.shutdown();
  return 0;
}
