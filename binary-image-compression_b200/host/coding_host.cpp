// golomb_encode / golomb_decode of the shim: thin calls into the device coder (include/bic_b200.h).
#include <cstdlib>
#include <iostream>

#include "GolombCoder.h"
#include "bic_b200.h"

static void ck(bic_status st, const char* what) {
  if (st == BIC_OK) return;
  std::cerr << "binary-image-compression_b200: " << what << ": " << bic_status_string(st) << std::endl;
  std::abort();
}

unsigned long golomb_encode(const binary_matrix& M, std::vector<uint8_t>& bytes, std::vector<uint64_t>* chunk_index,
                            unsigned long* nsamples) {
  bic_ctx* ctx = bic_host_context();
  bic_stream* s = 0;
  ck(bic_stream_create(ctx, &s), "stream_create");
  ck(bic_golomb_encode(ctx, M.device(), 256, s), "golomb_encode");
  bic_stream_info info;
  ck(bic_stream_get_info(s, &info), "stream_info");
  bytes.assign((info.bitcount + 7) / 8, 0);
  std::vector<uint64_t> idx(2 * info.nchunks + 2);
  ck(bic_stream_download(ctx, s, bytes.data(), bytes.size(), idx.data(), info.nchunks), "stream_download");
  idx.resize(2 * info.nchunks);
  if (chunk_index) *chunk_index = idx;
  if (nsamples) *nsamples = info.nsamples;
  ck(bic_stream_destroy(ctx, s), "stream_destroy");
  return info.bitcount;
}

void golomb_decode(const std::vector<uint8_t>& bytes, unsigned long bitcount, unsigned long nsamples,
                   const std::vector<uint64_t>& chunk_index, binary_matrix& M) {
  bic_ctx* ctx = bic_host_context();
  bic_stream* s = 0;
  ck(bic_stream_create(ctx, &s), "stream_create");
  bic_stream_info info;
  info.coder = BIC_CODER_GOLOMB;
  info.chunk_samples = 256;
  info.rows = M.get_rows();
  info.cols = M.get_cols();
  info.bitcount = bitcount;
  info.nchunks = chunk_index.size() / 2;
  info.nsamples = nsamples;
  ck(bic_stream_upload(ctx, s, &info, bytes.data(), chunk_index.data()), "stream_upload");
  ck(bic_golomb_decode(ctx, s, M.device()), "golomb_decode");
  M.device_written();
  ck(bic_stream_destroy(ctx, s), "stream_destroy");
}
