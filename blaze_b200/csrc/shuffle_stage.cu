// ShuffleWriteStage: the terminal stage of a plan rooted at ShuffleWriterExecNode.
//
// Reference behaviour restated (paths relative to /root/reference/native-engine/datafusion-ext-plans/src/):
//   ShuffleWriterExec::execute            shuffle_writer_exec.rs:109-165   (repartitioner by partitioning kind; empty output stream)
//   SortShuffleRepartitioner              shuffle/sort_repartitioner.rs:121-185 (insert_batch -> BufferedData; shuffle_write: .data + .index)
//   BufferedData::write                   shuffle/buffered_data.rs:123-158 (per partition: batches -> IpcCompressionWriter, finish_current_buf)
//   SingleShuffleRepartitioner            shuffle/single_repartitioner.rs:66-99
//   IpcCompressionWriter                  common/ipc_compression.rs:34-112 (blocks of u32 LE length ‖ LZ4 frame)
// GPU side (kernels_shuffle.cu): every pushed batch ("chunk") is partitioned and encoded into the uncompressed
// batch_serde bytes of each partition in ONE device buffer; the chunks' bytes are brought to the host, and finish()
// frames them per partition into compression blocks (host threads, lz4_frame.cc) and writes the two files exactly as
// the no-spill branch of shuffle_write does.  A partition of the file = the blocks of chunk 0, chunk 1, ... for it —
// the same shape the reference produces when it merges spills (sort_repartitioner.rs:226-246).
// Not on the GPU path (B200Q_ERR_UNSUPPORTED -> the host keeps its CPU operator): range and
// round-robin partitioning (they need the sort operator first, shuffle_writer_exec.rs:133-158), Binary / nested columns,
// more than 4096 partitions, codec zstd.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <thread>

#include "kernels_shuffle.cuh"
#include "lz4_frame.h"
#include "runtime.h"

namespace b200q {

namespace {

struct ShuffleChunk {
  int64_t rows = 0;
  std::vector<unsigned long long> part_off, part_rows;      // host copies: P + 1 byte offsets, P row counts
  std::vector<uint8_t> host;                                // encoded bytes (shuffle_output_to_host)
  DevMemP dev;                                              // encoded bytes in HBM (kept when the result stays on the device)
};

class ShuffleWriteStage : public Stage, public ShuffleResult {
  ShufSpec base_{};
  std::vector<int> hash_cols_;
  int P_ = 1;
  bool any_bits_ = false;
  std::string data_file_, index_file_;
  std::vector<ShuffleChunk> chunks_;
  std::vector<uint64_t> file_offsets_;
  DevMemP d_small_;                                         // counts | part_off | cursors

 public:
  ShuffleWriteStage(OpContext& cx, const SchemaDef& in, const PlanNode& node) {
    in_schema = in; out_schema = in;                        // ShuffleWriterExec::schema() = input schema (shuffle_writer_exec.rs:76-78); the stream is empty
    P_ = node.shuffle_kind == SHUFFLE_SINGLE ? 1 : (int)node.num_partitions;
    if (node.shuffle_kind == SHUFFLE_ROUND_ROBIN && P_ > 1) throw PlanError(B200Q_ERR_UNSUPPORTED, "round-robin shuffle partitioning sorts its input first (shuffle_writer_exec.rs:133-158): not on the GPU path");
    if (node.shuffle_kind == SHUFFLE_RANGE && P_ > 1) throw PlanError(B200Q_ERR_UNSUPPORTED, "range shuffle partitioning is not on the GPU path");
    if (P_ < 1) throw PlanError(B200Q_ERR_INVALID_PLAN, "shuffle writer with zero output partitions");
    if (P_ > SHUF_MAX_PARTS) throw PlanError(B200Q_ERR_UNSUPPORTED, "more than 4096 shuffle partitions");
    if (in.fields.size() > (size_t)SHUF_MAX_COLS) throw PlanError(B200Q_ERR_UNSUPPORTED, "more than 32 columns in a shuffled batch");
    data_file_ = node.data_file; index_file_ = node.index_file;
    base_.ncols = (int)in.fields.size(); base_.num_partitions = P_;
    base_.batch_size = cx.conf.batch_size > 0 ? cx.conf.batch_size : 10000;
    uint32_t k8 = 0, kw = 0;
    for (size_t i = 0; i < in.fields.size(); i++) {
      const FieldDef& f = in.fields[i];
      if (f.type.id == T_BINARY || f.type.id == T_NULL) throw PlanError(B200Q_ERR_UNSUPPORTED, "shuffle of a " + f.type.str() + " column is not on the GPU path");
      ShufCol& c = base_.col[i];
      c.width = (uint8_t)f.type.byte_width(); c.nullable = f.nullable ? 1 : 0; c.k8 = k8; c.kw = kw;
      if (c.nullable) k8++;
      if (c.width == 0) k8++; else kw += c.width;
      any_bits_ = any_bits_ || c.nullable || c.width == 0;
      used_input_cols.push_back((int)i);
    }
    base_.tot_k8 = k8; base_.tot_kw = kw;
    if (P_ > 1) {
      if (node.hash_exprs.empty()) throw PlanError(B200Q_ERR_INVALID_PLAN, "hash repartition without expressions");
      if (node.hash_exprs.size() > 8) throw PlanError(B200Q_ERR_UNSUPPORTED, "more than 8 hash partitioning expressions");
      for (auto& e : node.hash_exprs) {
        if (e->kind != E_COLUMN) throw PlanError(B200Q_ERR_UNSUPPORTED, "hash partitioning on a computed expression (project it first)");
        base_.key_col[base_.nkeys] = (int8_t)e->col_index; base_.key_phys[base_.nkeys] = (uint8_t)phys_of(e->type); base_.nkeys++;
      }
    }
    d_small_ = DevMem::alloc((size_t)(3 * P_ + 1) * 8, cx.stream);
  }

  void push(OpContext& cx, DevBatch& in, std::vector<DevBatch>&) override {
    const int64_t step = std::max<int64_t>(1, std::min<int64_t>(cx.conf.max_launch_rows > 0 ? cx.conf.max_launch_rows : (1LL << 27), 1LL << 27));
    for (int64_t r0 = 0; r0 < in.num_rows; r0 += step) encode_chunk(cx, in, r0, std::min(step, in.num_rows - r0));
  }

  void encode_chunk(OpContext& cx, DevBatch& in, int64_t r0, int64_t n) {
    ShufSpec sp = base_;
    for (int i = 0; i < sp.ncols; i++) {
      const DevColumn& dc = in.cols[i]; ShufCol& c = sp.col[i];
      const int64_t off = dc.offset + r0;
      if (off > 0xFFFFFFFFLL) throw ExecError(B200Q_ERR_UNSUPPORTED, "column offset beyond 2^32 rows");
      if (!dc.values) throw ExecError(B200Q_ERR_INVALID_ARG, "shuffle: column without a values buffer");
      c.values = c.width ? (const uint8_t*)dc.values->ptr + (size_t)off * c.width : (const uint8_t*)dc.values->ptr;
      c.validity = dc.validity ? (const uint8_t*)dc.validity->ptr : nullptr;
      c.bit_offset = (uint32_t)off;
      if (dc.validity && !c.nullable) throw ExecError(B200Q_ERR_INVALID_ARG, "shuffle: validity bitmap on a column the schema declares non-nullable");
    }
    unsigned long long* d_counts = (unsigned long long*)d_small_->ptr;
    unsigned long long* d_part_off = d_counts + P_;
    unsigned long long* d_cursors = d_part_off + P_ + 1;
    // every partition adds at most one short record per chunk: an upper bound of the encoded size that needs no host round trip
    const unsigned long long nrec_max = (unsigned long long)n / (unsigned long long)sp.batch_size + (unsigned long long)P_;
    const unsigned long long cap = (unsigned long long)sp.tot_kw * (unsigned long long)n + (unsigned long long)sp.tot_k8 * ((unsigned long long)n / 8 + nrec_max) +
                                   (unsigned long long)(5 + sp.ncols) * nrec_max + 64;
    DevMemP d_out = DevMem::alloc((size_t)cap, cx.stream);
    DevMemP d_pids = P_ > 1 ? DevMem::alloc((size_t)n * 2 + 16, cx.stream) : nullptr;
    B200Q_CUDA(cudaEventRecord(cx.ev0, cx.stream));
    B200Q_CUDA(cudaMemsetAsync(d_counts, 0, (size_t)P_ * 8, cx.stream));
    if (P_ > 1) cx.m.launches += launch_shuffle_pids(sp, n, (uint16_t*)d_pids->ptr, d_counts, cx.stream);
    else { const unsigned long long nn = (unsigned long long)n; B200Q_CUDA(cudaMemcpyAsync(d_counts, &nn, 8, cudaMemcpyHostToDevice, cx.stream)); }
    if (any_bits_) B200Q_CUDA(cudaMemsetAsync(d_out->ptr, 0, (size_t)cap, cx.stream));      // bit regions are OR-ed into
    cx.m.launches += launch_shuffle_layout(sp, d_counts, d_part_off, d_cursors, (uint8_t*)d_out->ptr, cx.stream);
    cx.m.launches += launch_shuffle_encode(sp, P_ > 1 ? (const uint16_t*)d_pids->ptr : nullptr, n, d_counts, d_part_off, d_cursors, (uint8_t*)d_out->ptr, cx.stream);
    cx.m.fast_launches++;
    B200Q_CUDA(cudaEventRecord(cx.ev1, cx.stream));
    ShuffleChunk ch; ch.rows = n; ch.part_off.resize((size_t)P_ + 1); ch.part_rows.resize((size_t)P_);
    B200Q_CUDA(cudaMemcpyAsync(ch.part_rows.data(), d_counts, (size_t)P_ * 8, cudaMemcpyDeviceToHost, cx.stream));
    B200Q_CUDA(cudaMemcpyAsync(ch.part_off.data(), d_part_off, (size_t)(P_ + 1) * 8, cudaMemcpyDeviceToHost, cx.stream));
    B200Q_CUDA(cudaStreamSynchronize(cx.stream));
    { float ms = 0; B200Q_CUDA(cudaEventElapsedTime(&ms, cx.ev0, cx.ev1)); cx.m.gpu_ms += ms; if (cx.cur_stage == 0) { cx.m.hot_ms += ms; cx.m.hot_rows += n; cx.m.hot_launches++; } }
    const unsigned long long total = ch.part_off[(size_t)P_];
    if (total > cap) throw ExecError(B200Q_ERR_EXECUTION, "internal: encoded shuffle chunk larger than its bound");
    if (cx.conf.shuffle_output_on_device) ch.dev = d_out;
    else {
      ch.host.resize((size_t)total);
      if (total) B200Q_CUDA(cudaMemcpyAsync(ch.host.data(), d_out->ptr, (size_t)total, cudaMemcpyDeviceToHost, cx.stream));
      B200Q_CUDA(cudaStreamSynchronize(cx.stream));
      cx.m.d2h_bytes += (int64_t)total;
    }
    chunks_.push_back(std::move(ch));
  }

  // ---- finish: frame + write (sort_repartitioner.rs:151-185, buffered_data.rs:123-158) ---------------------------------
  // one compression block = whole records of one (chunk, partition) up to ~4 MiB of payload (ipc_compression.rs:77-83 cuts
  // on 0.9 x 4 MiB of *compressed* bytes; where a block ends is not observable by a reader, :129-165)
  void compress_partition(int p, std::vector<uint8_t>& out) const {
    const unsigned long long B = (unsigned long long)base_.batch_size, F = shuf_record_bytes(base_, B);
    constexpr unsigned long long TARGET = 4ull << 20;
    for (const ShuffleChunk& ch : chunks_) {
      const unsigned long long t = ch.part_rows[(size_t)p];
      if (t == 0) continue;
      const uint8_t* src = ch.host.data() + ch.part_off[(size_t)p];
      const unsigned long long len = ch.part_off[(size_t)p + 1] - ch.part_off[(size_t)p];
      const unsigned long long per_block = std::max<unsigned long long>(1, TARGET / F) * F;     // whole records
      for (unsigned long long pos = 0; pos < len; pos += per_block) {
        const unsigned long long blen = std::min(per_block, len - pos);
        const size_t at = out.size();
        out.resize(at + 4);
        lz4_frame_append(src + pos, (size_t)blen, out);
        const uint32_t framed = (uint32_t)(out.size() - at - 4);
        memcpy(out.data() + at, &framed, 4);
      }
    }
  }

  void finish(OpContext& cx, std::vector<DevBatch>&) override {
    if (cx.conf.shuffle_output_on_device || data_file_.empty()) return;
    std::vector<std::vector<uint8_t>> comp((size_t)P_);
    std::atomic<int> next{0};
    const unsigned hw = std::thread::hardware_concurrency();
    const int nthreads = (int)std::max(1u, std::min(std::min(hw ? hw : 4u, 32u), (unsigned)P_));
    auto work = [&] { for (int p = next.fetch_add(1); p < P_; p = next.fetch_add(1)) compress_partition(p, comp[(size_t)p]); };
    std::vector<std::thread> th;
    for (int i = 1; i < nthreads; i++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    FILE* fd = fopen(data_file_.c_str(), "wb");
    if (!fd) throw ExecError(B200Q_ERR_EXECUTION, "shuffle write error: cannot open " + data_file_);
    file_offsets_.assign((size_t)P_ + 1, 0);
    uint64_t pos = 0; bool ok = true;
    for (int p = 0; p < P_; p++) {
      file_offsets_[(size_t)p] = pos;
      if (!comp[(size_t)p].empty()) ok = ok && fwrite(comp[(size_t)p].data(), 1, comp[(size_t)p].size(), fd) == comp[(size_t)p].size();
      pos += comp[(size_t)p].size();
    }
    file_offsets_[(size_t)P_] = pos;
    ok = (fclose(fd) == 0) && ok;
    FILE* fi = fopen(index_file_.c_str(), "wb");
    if (!fi) throw ExecError(B200Q_ERR_EXECUTION, "shuffle write error: cannot open " + index_file_);
    for (uint64_t o : file_offsets_) { const int64_t v = (int64_t)o; ok = ok && fwrite(&v, 8, 1, fi) == 1; }     // little-endian i64 (sort_repartitioner.rs:181-185)
    ok = (fclose(fi) == 0) && ok;
    if (!ok) throw ExecError(B200Q_ERR_EXECUTION, "shuffle write error: short write");
  }

  // ---- ShuffleResult ------------------------------------------------------------------------------------------------
  int64_t chunk_count() const override { return (int64_t)chunks_.size(); }
  void chunk(int64_t i, b200q_shuffle_chunk* out) const override {
    const ShuffleChunk& ch = chunks_.at((size_t)i);
    out->num_partitions = P_; out->rows = ch.rows;
    out->on_device = ch.dev ? 1 : 0;
    out->data = ch.dev ? (const uint8_t*)ch.dev->ptr : ch.host.data();
    out->part_off = (const uint64_t*)ch.part_off.data(); out->part_rows = (const uint64_t*)ch.part_rows.data();
  }
};

}  // namespace

std::unique_ptr<Stage> make_shuffle_write_stage(OpContext& cx, const SchemaDef& in_schema, const PlanNode& node) {
  return std::unique_ptr<Stage>(new ShuffleWriteStage(cx, in_schema, node));
}

}  // namespace b200q
