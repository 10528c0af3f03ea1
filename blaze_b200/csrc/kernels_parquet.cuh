// Descriptors and launchers of the Parquet column-chunk decode kernels (kernels_parquet.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"

namespace b200q {

// a run of consecutive entries of one column chunk; `start` counts rows (level runs) or stored values (value runs)
enum PqRunKind : uint8_t { PQR_RLE = 0 /* one value repeated */, PQR_BITPACKED = 1 /* bw-bit values from bit `off` */, PQR_PLAIN = 2 /* fixed-width values from byte `off` */ };
struct PqDevRun { uint32_t start, count; uint8_t kind, bw; uint8_t _pad[6]; unsigned long long off_or_value; };

enum PqOut : uint8_t { PQO_I8 = 0, PQO_I16, PQO_I32, PQO_I64, PQO_DEC_FROM_I32, PQO_DEC_FROM_I64, PQO_DEC_FROM_FLBA, PQO_BOOL_BYTES };
struct PqDecodeSpec {
  const uint8_t* bytes;               // the chunk's decompressed page bodies, back to back
  const PqDevRun* value_runs; int32_t n_value_runs;
  const uint8_t* dict;                // PLAIN dictionary entries (src_width bytes each), null when no page is dictionary-encoded
  int32_t dict_count;
  int32_t src_width;                  // bytes per stored value (4, 8, FLBA length; 0: Boolean bits)
  uint8_t out_kind; uint8_t _pad[3];
};

// valid[r] = definition level of row r (0 / 1) from level runs
int launch_pq_levels(const uint8_t* bytes, const PqDevRun* runs, int n_runs, int64_t n_rows, uint8_t* valid, cudaStream_t s);
// out[r] = decoded value of row r (ordinal[r]-th stored value when valid, 0 otherwise); ordinal == null: every row is stored (ordinal = r)
int launch_pq_decode(const PqDecodeSpec& sp, const uint8_t* valid, const int32_t* ordinal, int64_t n_rows, void* out, int* d_err, cudaStream_t s);

}  // namespace b200q
