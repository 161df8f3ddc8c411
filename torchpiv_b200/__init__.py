"""torchpiv_b200 -- B200-native (sm_100a) implementation of TorchPIV's PIV cross-correlation
hot path behind the reference's own Python API (see backend.py)."""
from .backend import (DeviceMap, IterModMap, OfflinePIV, PIVDataset, ToTensor,  # noqa: F401
                      biliniar_interpolation_CWS, correalte_fft, correlation_to_displacement,
                      extended_search_area_piv, get_coordinates, get_field_shape,
                      interpolation_DWS, moving_window_array, natural_keys, piv_iteration_CWS,
                      piv_iteration_DWS)
from .engine import PIVPlan  # noqa: F401

__version__ = "0.1.0"
