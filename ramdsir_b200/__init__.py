"""ramdsir_b200 -- B200-native Random Amplitude Mixup (the data-parallel hot path of RAM-DSIR).

Public surface
  per-sample drop-ins (reference signatures):  extract_amp_spectrum, low_freq_mutate_np, source_to_target_freq
  batched entry points:                        build_amp_bank, ram_mix_batch, MixCall (a validated, argument-bound launch)
  C ABI:                                       include/ramdsir.h -> ramdsir_b200/libramdsir.so
"""
from .api import (FUNDUS_EPILOGUE, NO_EPILOGUE, PROSTATE_EPILOGUE, MixCall, band_halfwidth, build_amp_bank, kernel_path,
                  prepare, ram_mix_batch, u8_supported, workspace_bytes)
from .freq import extract_amp_spectrum, low_freq_mutate_np, source_to_target_freq

__all__ = ["extract_amp_spectrum", "low_freq_mutate_np", "source_to_target_freq", "build_amp_bank", "u8_supported",
           "ram_mix_batch", "MixCall", "band_halfwidth", "kernel_path", "workspace_bytes", "prepare",
           "FUNDUS_EPILOGUE", "PROSTATE_EPILOGUE", "NO_EPILOGUE"]
__version__ = "0.1.0"
