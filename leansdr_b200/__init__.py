"""leansdr_b200 -- B200-native (sm_100a) DVB-S receive path behind leansdr's runnable API.

The package is a thin loader around libleandvb_b200.so (hand-written CUDA
kernels + the C ABI declared in include/leandvb_b200.h).  There is no CPU
fallback: importing works anywhere, but creating a Receiver needs the built
library and a B200.
"""
from .capi import (Config, LdvbError, Meas, Receiver, default_config, deint_rs, fir_cf32, host_table,  # noqa: F401
                   host_register, host_unregister, ring_unique_id, load, rs_decode, EXPORTS, LIB_PATH, RX_EXACT, RX_FAST)

from .tx import TX_EXPORTS, Transmitter, TxConfig, fir_resampler_cf32, host_taps, tx_config  # noqa: F401

__all__ = ["TX_EXPORTS", "Transmitter", "TxConfig", "fir_resampler_cf32", "host_taps", "tx_config", "Config", "LdvbError", "Meas", "Receiver", "default_config", "deint_rs", "fir_cf32", "host_table", "host_register", "host_unregister", "ring_unique_id",
           "load", "rs_decode", "EXPORTS", "LIB_PATH", "RX_EXACT", "RX_FAST"]
