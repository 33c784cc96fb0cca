// MSB-first bit writer/reader: the BinaryFileWriter/BinaryFileReader the reference's coders were
// written against (src/GolombCoder.h:14,24-27; src/GolombDecoder.h:13) but which it does not ship.
// Stream bit t is in byte t/8 at mask 0x80 >> (t % 8).
#ifndef BIC_HOST_BITIO_H
#define BIC_HOST_BITIO_H
#include <cstdint>
#include <vector>
class BinaryFileWriter {
 public:
  void writeBits(unsigned value, unsigned nbits) { for (unsigned i = nbits; i-- > 0;) put((value >> i) & 1u); }
  void writeZeros(unsigned long n) { for (unsigned long i = 0; i < n; ++i) put(0); }
  unsigned long bits() const { return nbits_; }
  const std::vector<uint8_t>& bytes() const { return buf_; }
 private:
  void put(unsigned b) {
    if ((nbits_ & 7) == 0) buf_.push_back(0);
    if (b) buf_.back() |= (uint8_t)(0x80u >> (nbits_ & 7));
    nbits_++;
  }
  std::vector<uint8_t> buf_;
  unsigned long nbits_ = 0;
};
class BinaryFileReader {
 public:
  BinaryFileReader(const uint8_t* p, unsigned long nbits) : p_(p), n_(nbits), pos_(0) {}
  unsigned readBits(unsigned nbits) { unsigned v = 0; for (unsigned i = 0; i < nbits; ++i) v = (v << 1) | get(); return v; }
  unsigned countZeros() { unsigned z = 0; while (pos_ < n_ && !peek()) { z++; pos_++; } return z; }
  unsigned long position() const { return pos_; }
 private:
  unsigned peek() const { return pos_ < n_ ? (p_[pos_ >> 3] >> (7 - (pos_ & 7))) & 1u : 1u; }
  unsigned get() { const unsigned b = peek(); pos_++; return b; }
  const uint8_t* p_;
  unsigned long n_, pos_;
};
#endif
