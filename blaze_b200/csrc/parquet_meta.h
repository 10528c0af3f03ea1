// Parquet file metadata + page framing, host side (parquet_meta.cc): Thrift compact protocol reader for FileMetaData /
// PageHeader, page decompression (Snappy, uncompressed), RLE / bit-packed hybrid run tables.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ir.h"

namespace b200q {

enum PqType : int { PQ_BOOLEAN = 0, PQ_INT32 = 1, PQ_INT64 = 2, PQ_INT96 = 3, PQ_FLOAT = 4, PQ_DOUBLE = 5, PQ_BYTE_ARRAY = 6, PQ_FIXED_LEN_BYTE_ARRAY = 7 };
enum PqCodec : int { PQ_UNCOMPRESSED = 0, PQ_SNAPPY = 1, PQ_GZIP = 2, PQ_LZO = 3, PQ_BROTLI = 4, PQ_LZ4 = 5, PQ_ZSTD = 6, PQ_LZ4_RAW = 7 };
enum PqEncoding : int { PQ_PLAIN = 0, PQ_PLAIN_DICTIONARY = 2, PQ_RLE = 3, PQ_BIT_PACKED = 4, PQ_RLE_DICTIONARY = 8 };
enum PqPageType : int { PQ_DATA_PAGE = 0, PQ_INDEX_PAGE = 1, PQ_DICTIONARY_PAGE = 2, PQ_DATA_PAGE_V2 = 3 };

struct PqColumnSchema {                 // a leaf of a FLAT schema (required / optional primitive)
  std::string name;
  int type = 0, type_length = 0;
  bool optional = false;
  int converted_type = -1, scale = 0, precision = 0;
  bool logical_date = false, logical_ts_micros = false, logical_decimal = false;
  int int_bits = 0; bool int_signed = true;   // LogicalType INTEGER
  DType arrow;                          // the Arrow type this repo maps the column to (T_NULL: not on the GPU path)
};

struct PqStats { bool has_min = false, has_max = false; std::string min, max; int64_t null_count = -1; };

struct PqColumnChunk {
  int type = 0, codec = 0;
  int64_t num_values = 0, total_compressed_size = 0, total_uncompressed_size = 0;
  int64_t data_page_offset = 0, dictionary_page_offset = -1;
  PqStats stats;
  int64_t start() const { return dictionary_page_offset > 0 && dictionary_page_offset < data_page_offset ? dictionary_page_offset : data_page_offset; }
};

struct PqRowGroup { int64_t num_rows = 0; std::vector<PqColumnChunk> columns; };

struct PqFileMeta {
  int64_t num_rows = 0;
  std::vector<PqColumnSchema> columns;  // leaves, in file order
  std::vector<PqRowGroup> row_groups;
  bool flat = true;                     // false: nested / repeated fields present (not on the GPU path)
};

// footer = the Thrift-encoded FileMetaData (without the trailing length + magic)
PqFileMeta parquet_parse_footer(const uint8_t* footer, size_t n);

struct PqRun { uint32_t count; uint32_t is_rle; uint64_t value_or_bit_offset; };      // RLE run: the value; bit-packed run: bit offset of its first value inside the page's value bytes

// growable byte buffer for decompressed page bodies.  Plain heap by default; the scan backs it with pinned host memory (alloc_fn / free_fn) so that
// the upload of a column chunk is a DMA straight out of the buffer the pages were decompressed into.  Bytes appended are not initialised.
struct ByteBuf {
  uint8_t* p = nullptr; size_t n = 0, cap = 0;
  void* (*alloc_fn)(size_t) = nullptr; void (*free_fn)(void*) = nullptr;
  ByteBuf() = default;
  ByteBuf(const ByteBuf&) = delete; ByteBuf& operator=(const ByteBuf&) = delete;
  ~ByteBuf() { release(); }
  void release() { if (p) { if (free_fn) free_fn(p); else free(p); } p = nullptr; n = cap = 0; }
  void reserve(size_t want);
  uint8_t* grow(size_t add) { reserve(n + add); uint8_t* r = p + n; n += add; return r; }
  void append(const uint8_t* s, size_t len) { if (len) memcpy(grow(len), s, len); }
  void clear() { n = 0; }
  const uint8_t* data() const { return p; }
  size_t size() const { return n; }
};

// one decoded (decompressed) page, described for the device
struct PqPage {
  int type = 0;                         // PQ_DATA_PAGE / PQ_DATA_PAGE_V2 / PQ_DICTIONARY_PAGE
  int32_t num_values = 0;               // entries incl. NULLs (data pages); dictionary entries
  int encoding = 0;
  size_t base = 0, size = 0;            // the decompressed page body = out[base, base + size) of the buffer handed to parquet_read_pages
  size_t values_offset = 0;             // where the value bytes start inside the body
  std::vector<PqRun> def_runs;          // definition levels (max level 1) as runs; empty: every value is present
  int dict_bit_width = 0;
  std::vector<PqRun> idx_runs;          // dictionary indices as runs (offsets relative to values_offset + 1)
  int64_t non_null = 0;                 // values actually stored
};

// walks the pages of one column chunk (`chunk` = its bytes [start, start + total_compressed_size)); throws PlanError(UNSUPPORTED) for
// codecs / encodings outside the GPU path
// data page bodies are appended to `out` (one contiguous buffer per chunk: what the device reads), the dictionary page's to `dict_out`;
// both keep their capacity across calls so that a scan does not re-fault its buffers for every row group
std::vector<PqPage> parquet_read_pages(const uint8_t* chunk, size_t n, const PqColumnChunk& cc, const PqColumnSchema& cs, ByteBuf& out, ByteBuf& dict_out);

size_t snappy_uncompress(const uint8_t* src, size_t n, ByteBuf& out);     // raw Snappy block format; appends to `out`, returns the decompressed size

}  // namespace b200q
