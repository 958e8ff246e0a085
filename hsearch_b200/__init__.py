"""hsearch_b200 -- B200-native (sm_100a) implementation of the HSEARCH hot path.

The product is libhsearch_b200.so (hand-written CUDA behind the C ABI of
include/hsearch_b200.h).  This package is the thin host-side mirror used by the
tests, the bench and the Python tooling; it contains no compute of its own.
"""
from .capi import (HIT_DTYPE, HS_FLAG_HASH_AUDIT, HS_FLAG_HASH_EXACT, HS_FLAG_SCALAR_FILTER, HS_FLAG_SORT_HITS,  # noqa: F401
                   HS_METRIC_BLOSUM_INT,
                   HS_METRIC_EUCLID_FP64, HS_PRED_D2_LE_R2, HS_PRED_SQRT_LE_R, HS_TABLE_FULL, HS_TABLE_PRINT6,
                   HsError)
from .index import (AA_ORDER, HSearch, blosum_metric, coordinates, encode, generate_projection,  # noqa: F401
                    pack_key_string)
