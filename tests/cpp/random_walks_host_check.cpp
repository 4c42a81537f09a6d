// random_walks_host_check.cpp -- TEST-ONLY: runs the walk arithmetic of TheaGpu::SurfaceFeatures::Local::RandomWalksT (LocalFeatures.hpp) on
// the host, with a stand-in surface that returns seed samples the test computed beforehand, so that the CPU suite can compare the C++
// caller with the oracle (tests/test_local_features.py).  On a GPU the surface is a GpuPointBvh; nothing else differs.
// Usage: random_walks_host_check in.bin out.bin
//   in : int64 P, K, Q, steps, walks, seed ; float pos[P*3] ; int32 graph[P*K] ; float q[Q*3] ; int32 seed_sample[Q]
//   out: double features[Q*steps*3]
#include "../../include/thea_gpu/LocalFeatures.hpp"

#include <cstdio>
#include <cstdlib>
#include <vector>

struct GivenSeeds
{
  std::vector<int32_t> seeds;
  void closestElementBatch(float const *, TheaGpu::intx Q, int32_t * idx, float *, float *) const { for (TheaGpu::intx i = 0; i < Q; ++i) idx[i] = seeds[(size_t)i]; }
};

template <typename T> static std::vector<T> rd(FILE * f, size_t n) { std::vector<T> v(n); if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } return v; }

int main(int argc, char ** argv)
{
  if (argc < 3) return 2;
  FILE * fi = fopen(argv[1], "rb"); if (!fi) return 2;
  std::vector<int64_t> h = rd<int64_t>(fi, 6);
  int64_t P = h[0], K = h[1], Q = h[2], steps = h[3], walks = h[4]; uint32_t seed = (uint32_t)h[5];
  std::vector<float> pos = rd<float>(fi, (size_t)(3 * P)); std::vector<int32_t> graph = rd<int32_t>(fi, (size_t)(P * K));
  std::vector<float> q = rd<float>(fi, (size_t)(3 * Q));
  GivenSeeds surf; surf.seeds = rd<int32_t>(fi, (size_t)Q);
  fclose(fi);
  TheaGpu::SurfaceFeatures::Local::TheaRandom rng(seed);
  TheaGpu::SurfaceFeatures::Local::RandomWalksT<TheaGpu::Vec3, GivenSeeds> rw(&surf, pos.data(), graph.data(), K, &rng);
  std::vector<double> out((size_t)(3 * steps * Q));
  rw.computeBatch(q.data(), Q, steps, out.data(), walks);
  FILE * fo = fopen(argv[2], "wb"); if (!fo) return 2;
  fwrite(out.data(), sizeof(double), out.size(), fo); fclose(fo);
  return 0;
}
