// Stand-in for src/GolombCoder.h:19-28 (+ the writer its commented lines call, :22-25) and for the
// uncompilable src/GolombDecoder.h:19-26 (unsigned samples). Serial, host side: the bit-exact spec the
// device encoder (bic_golomb_encode) is tested against, and the drop-in for callers that code a few
// samples at a time (compress*_test.cpp). For whole matrices use golomb_encode() below.
#ifndef BIC_HOST_GOLOMBCODER_H
#define BIC_HOST_GOLOMBCODER_H
#include <cstdint>
#include <vector>
#include "BitIO.h"
#include "Golomb.h"
#include "binmat.h"

class GolombCoder : public Golomb {
 public:
  GolombCoder() : Golomb(), bitcount(0), file(0) {}
  explicit GolombCoder(BinaryFileWriter* f) : Golomb(), bitcount(0), file(f) {}
  void codeSample(unsigned sample) {  // src/GolombCoder.cpp:29-34
    binaryEncode(sample, k);
    adapt(sample);
  }
  long bitcount;
 private:
  BinaryFileWriter* file;
  void binaryEncode(unsigned sample, unsigned kk) {  // src/GolombCoder.cpp:13-27
    const unsigned unary = sample >> kk;
    if (file) {
      file->writeBits(kk ? (sample & (0xFFFFFFFFu >> (32 - kk))) : 0u, kk);  // :22
      file->writeZeros(unary);                                               // :24
      file->writeBits(1, 1);                                                 // :25
    }
    bitcount += kk + unary + 1;
  }
};

class GolombDecoder : public Golomb {
 public:
  explicit GolombDecoder(BinaryFileReader* f) : Golomb(), file(f) {}
  unsigned decodeSample() {  // src/GolombDecoder.cpp:15-40, unsigned path
    const unsigned binary = file->readBits(k);
    const unsigned unary = file->countZeros();
    file->readBits(1);
    const unsigned sample = (unary << k) | binary;
    adapt(sample);
    return sample;
  }
 private:
  BinaryFileReader* file;
};

// Whole-matrix coding on the device: zero-run lengths of M read row-major, a virtual one closing the
// last run; byte-identical to feeding those runs to GolombCoder(BinaryFileWriter*). Returns bitcount.
unsigned long golomb_encode(const binary_matrix& M, std::vector<uint8_t>& bytes, std::vector<uint64_t>* chunk_index = 0,
                            unsigned long* nsamples = 0);
// chunk_index / nsamples as returned by golomb_encode (the stream is not self-synchronising)
void golomb_decode(const std::vector<uint8_t>& bytes, unsigned long bitcount, unsigned long nsamples,
                   const std::vector<uint64_t>& chunk_index, binary_matrix& M);
#endif
