// Stand-in for src/eg.h:6-38 / src/eg.cpp:2-37 (the "DUMMY Golomb Runlength coder").
#ifndef BIC_HOST_EG_H
#define BIC_HOST_EG_H
#include "BitIO.h"
class EG {
 public:
  EG() { g = 1; blockSize = 1; lutIndex = 0; }
  unsigned g;
  unsigned blockSize;
  int lutIndex;
  void incBlockSize() { if (lutIndex < 31) lutIndex++; g = lut(lutIndex); blockSize = 1u << g; }
  void decBlockSize() { if (lutIndex > 0) lutIndex--; g = lut(lutIndex); blockSize = 1u << g; }
 private:
  static unsigned lut(int i) {  // src/eg.cpp:2
    static const short t[32] = {0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 9, 10, 11, 12, 13, 14, 15};
    return (unsigned)t[i];
  }
};
class EGCoder : public EG {
 public:
  EGCoder() : EG(), bitcount(0), file(0) {}
  explicit EGCoder(BinaryFileWriter* f) : EG(), bitcount(0), file(f) {}
  void codeRun(int len, bool eol) {  // src/eg.cpp:20-37 (block growth disabled there, :25)
    while ((unsigned)len >= blockSize) {
      len -= (int)blockSize;
      if (file) file->writeBits(1, 1);
      bitcount++;
    }
    if (eol) {
      if (file) file->writeBits(1, 1);
      bitcount++;
    } else {
      if (file) { file->writeBits(0, 1); file->writeBits((unsigned)len, g); }
      bitcount += (g + 1);
      decBlockSize();
    }
  }
  unsigned long bitcount;
 private:
  BinaryFileWriter* file;
};
#endif
