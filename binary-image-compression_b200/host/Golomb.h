// Stand-in for src/Golomb.h:12-29: the adaptive Golomb-Rice state.
#ifndef BIC_HOST_GOLOMB_H
#define BIC_HOST_GOLOMB_H
class Golomb {
 public:
  Golomb() : accumulatedError(0), samples(0), k(1) {}
 protected:
  unsigned accumulatedError;
  unsigned samples;
  unsigned k;
  // k = min{k : (samples << k) >= accumulatedError} in unsigned arithmetic (src/GolombCoder.cpp:33);
  // the search stops at 31 (k >= 32 trips the reference's assert, src/GolombCoder.cpp:14)
  void adapt(unsigned sample) {
    samples++;
    accumulatedError += sample;
    for (k = 0; k < 31 && (samples << k) < accumulatedError; k++) {}
  }
};
#endif
