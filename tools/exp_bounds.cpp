// exp_bounds.cpp -- DEVELOPMENT AID (not product, not test): counts node visits / triangle tests of closest-point
// traversals under alternative bounding volumes, to pick the device layout before spending GPU time.
//   g++ -std=c++17 -O2 -ffp-contract=off -o /tmp/exp_bounds tools/exp_bounds.cpp && /tmp/exp_bounds 708 20000
#include "../tests/emu/emu.cpp"
#include <cstdio>
#include <cstdlib>
#include <functional>

struct Obb { v3 c, a0, a1, a2; float h0, h1, h2; };   // centre, axes (a0 = mean normal), half extents

static uint64_t sm64(uint64_t seed, uint64_t i)
{
  uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31);
}
static float u01(uint64_t seed, uint64_t i) { return (float)(sm64(seed, i) >> 40) * (1.0f / 16777216.0f); }

static std::vector<float> torus(int n)
{
  std::vector<float> V((size_t)n * n * 3);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j)
  {
    double u = 2 * M_PI * i / n, v = 2 * M_PI * j / n;
    double r = 0.4 * (1 + 0.15 * sin(3 * u) * cos(5 * v) + 0.05 * sin(11 * u + 2 * v));
    double x = (1 + r * cos(v)) * cos(u), y = (1 + r * cos(v)) * sin(u), z = r * sin(v);
    float * p = &V[((size_t)i * n + j) * 3]; p[0] = (float)x; p[1] = (float)y; p[2] = (float)z;
  }
  std::vector<float> T((size_t)n * n * 18);
  size_t k = 0;
  auto put = [&](int i, int j) { float * p = &V[((size_t)(i % n) * n + (j % n)) * 3]; T[k++] = p[0]; T[k++] = p[1]; T[k++] = p[2]; };
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j)
  { put(i, j); put(i + 1, j); put(i + 1, j + 1); put(i, j); put(i + 1, j + 1); put(i, j + 1); }
  return T;
}

int main(int argc, char ** argv)
{
  int grid = argc > 1 ? atoi(argv[1]) : 708; int Q = argc > 2 ? atoi(argv[2]) : 20000; int leaf = argc > 3 ? atoi(argv[3]) : 4;
  std::vector<float> T = torus(grid);
  Emu E; E.tri = true; E.n = (int)(T.size() / 9); E.leaf = leaf;
  build(E, T.data(), nullptr);
  int nn = (int)E.nodes.size();
  printf("tris %d nodes %d leaf %d\n", E.n, nn, leaf);
  // element range of every child
  std::vector<int> cf((size_t)nn * 2), cl((size_t)nn * 2);
  std::function<void(int, int &, int &)> rng = [&](int i, int & f, int & l) {
    f = 1 << 30; l = -1;
    for (int c = 0; c < 2; ++c)
    {
      int ch = f2i(c == 0 ? E.nodes[(size_t)i].q3.x : E.nodes[(size_t)i].q3.y); int a, b;
      if (ch < 0) { int code = ~ch; a = code >> TGQ_LEAF_BITS; b = a + (code & TGQ_LEAF_MASK) - 1; } else rng(ch, a, b);
      cf[(size_t)i * 2 + c] = a; cl[(size_t)i * 2 + c] = b; f = std::min(f, a); l = std::max(l, b);
    }
  };
  int rf, rl; rng(0, rf, rl);
  // OBB per child
  std::vector<Obb> obb((size_t)nn * 2);
  float pad = E.pad;
  for (int i = 0; i < nn; ++i) for (int c = 0; c < 2; ++c)
  {
    int a = cf[(size_t)i * 2 + c], b = cl[(size_t)i * 2 + c];
    Obb o;
    if (b < a) { o.c = mk3(0, 0, 0); o.a0 = mk3(1, 0, 0); o.a1 = mk3(0, 1, 0); o.a2 = mk3(0, 0, 1); o.h0 = o.h1 = o.h2 = -FLT_MAX; obb[(size_t)i * 2 + c] = o; continue; }
    v3 sum = mk3(0, 0, 0);
    for (int k = a; k <= b; ++k)
    {
      TriRec const & r = E.tris[(size_t)k]; v3 v0 = mk3(r.q0.x, r.q0.y, r.q0.z);
      sum = add3(sum, cross3(sub3(mk3(r.q1.x, r.q1.y, r.q1.z), v0), sub3(mk3(r.q2.x, r.q2.y, r.q2.z), v0)));
    }
    v3 d = normalized3(sum);
    if (!(sqn3(d) > 0.5f)) d = mk3(0, 0, 1);
    int ax = 0; float m = fabsf(d.x); if (fabsf(d.y) < m) { ax = 1; m = fabsf(d.y); } if (fabsf(d.z) < m) ax = 2;
    v3 e = mk3(ax == 0, ax == 1, ax == 2);
    v3 t1 = normalized3(sub3(e, mul3(d, dot3(d, e))));
    v3 t2 = cross3(d, t1);
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int k = a; k <= b; ++k)
    {
      TriRec const & r = E.tris[(size_t)k];
      v3 vs[3] = {mk3(r.q0.x, r.q0.y, r.q0.z), mk3(r.q1.x, r.q1.y, r.q1.z), mk3(r.q2.x, r.q2.y, r.q2.z)};
      for (v3 v : vs)
      {
        float s[3] = {dot3(d, v), dot3(t1, v), dot3(t2, v)};
        for (int j = 0; j < 3; ++j) { mn[j] = fminf(mn[j], s[j]); mx[j] = fmaxf(mx[j], s[j]); }
      }
    }
    float cs[3], hs[3];
    for (int j = 0; j < 3; ++j) { cs[j] = 0.5f * (mn[j] + mx[j]); hs[j] = 0.5f * (mx[j] - mn[j]) + 3 * pad; }
    o.a0 = d; o.a1 = t1; o.a2 = t2; o.c = add3(add3(mul3(d, cs[0]), mul3(t1, cs[1])), mul3(t2, cs[2]));
    o.h0 = hs[0]; o.h1 = hs[1]; o.h2 = hs[2];
    obb[(size_t)i * 2 + c] = o;
  }
  // queries: uniform in the x1.5 box
  v3 lo = mk3(FLT_MAX, FLT_MAX, FLT_MAX), hi = mk3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
  for (size_t i = 0; i < T.size() / 3; ++i) { v3 p = ld(T.data(), (int64_t)i); lo = min3(lo, p); hi = max3(hi, p); }
  v3 cc = mul3(add3(lo, hi), 0.5f), he = mul3(sub3(hi, lo), 0.75f);
  std::vector<v3> qs((size_t)Q);
  for (int i = 0; i < Q; ++i)
    qs[(size_t)i] = mk3(cc.x + (2 * u01(7, 3 * i) - 1) * he.x, cc.y + (2 * u01(7, 3 * i + 1) - 1) * he.y, cc.z + (2 * u01(7, 3 * i + 2) - 1) * he.z);

  BvhView V = E.view();
  auto obb_lb = [&](Obb const & o, v3 p, float mq) {
    v3 pc = sub3(p, o.c);
    float x = fmaxf(fabsf(dot3(o.a0, pc)) - o.h0 - mq, 0.0f), y = fmaxf(fabsf(dot3(o.a1, pc)) - o.h1 - mq, 0.0f), z = fmaxf(fabsf(dot3(o.a2, pc)) - o.h2 - mq, 0.0f);
    return (x * x + y * y + z * z) * TGQ_LB_DEFLATE;
  };
  // mode 0: AABB only, 1: AABB + pancake (current), 2: OBB only, 3: OBB + AABB ; screen 0: none, 1: pancake circle (current)
  for (int mode = 2; mode < 3; ++mode) for (int order = 0; order < 3; ++order) for (int tight = 0; tight < 1; ++tight)
  {
    long long nodes = 0, scr = 0, exact = 0, mism = 0;
    std::vector<int> pernode((size_t)Q);
    for (int qi = 0; qi < Q; ++qi)
    {
      v3 p = qs[(size_t)qi]; float mq = query_margin(p, E.pad);
      float ref = INFINITY; int ri = -1; v3 rcp;
      closest_query<true, true>(V, E.slabs.data(), E.pad, p, -1, ref, ri, rcp);
      float best = tight ? nextafterf(ref, INFINITY) : INFINITY; int bi = -1; v3 bcp = mk3(0, 0, 0);
      int stk[256]; float slb[256]; int sp = 0; int cur = 0; long long n0 = nodes;
      for (;;)
      {
        while (cur >= 0)
        {
          ++nodes;
          NodeRec nd = E.nodes[(size_t)cur]; v3 lo0, hi0, lo1, hi1; node_boxes(nd, lo0, hi0, lo1, hi1);
          float lb[2], key[2]; int ch[2] = {f2i(nd.q3.x), f2i(nd.q3.y)};
          for (int c = 0; c < 2; ++c)
          {
            float aabb = box_point_sqdist(c ? lo1 : lo0, c ? hi1 : hi0, p) * TGQ_LB_DEFLATE;
            float l = 0;
            if (mode == 0) l = aabb;
            else if (mode == 1)
            {
              SlabRec const & s = E.slabs[(size_t)cur];
              float pk = c == 0 ? pancake_sqdist(box_centre(lo0, hi0), mk3(s.a.x, s.a.y, s.a.z), s.a.w, s.b.x, s.c.z, p, mq)
                                : pancake_sqdist(box_centre(lo1, hi1), mk3(s.b.y, s.b.z, s.b.w), s.c.x, s.c.y, s.c.w, p, mq);
              l = fmaxf(aabb, pk);
            }
            else if (mode == 2) l = obb_lb(obb[(size_t)cur * 2 + c], p, mq);
            else l = fmaxf(aabb, obb_lb(obb[(size_t)cur * 2 + c], p, mq));
            lb[c] = l;
            Obb const & oo = obb[(size_t)cur * 2 + c];
            float dc = sqn3(sub3(p, oo.c));
            key[c] = order == 0 ? l : order == 1 ? dc : (sqrtf(l) + sqrtf(dc));
          }
          bool g0 = lb[0] <= best, g1 = lb[1] <= best;
          if (g0 && g1) { bool sw = key[1] < key[0]; stk[sp] = sw ? ch[0] : ch[1]; slb[sp] = sw ? lb[0] : lb[1]; ++sp; cur = sw ? ch[1] : ch[0]; }
          else if (g0) cur = ch[0]; else if (g1) cur = ch[1]; else goto pop;
        }
        {
          int code = ~cur; int first = code >> TGQ_LEAF_BITS, count = code & TGQ_LEAF_MASK;
          long long p0 = g_cnt_prim;
          scr += count;
          closest_leaf<true, true>(V, first, count, p, -1, mq, best, bi, bcp);
          exact += g_cnt_prim - p0;
        }
      pop:
        for (;;) { if (sp == 0) goto done; --sp; if (slb[sp] <= best) { cur = stk[sp]; break; } }
      }
    done:
      pernode[(size_t)qi] = (int)(nodes - n0);
      if (bi != ri || best != ref) ++mism;
    }
    std::sort(pernode.begin(), pernode.end());
    printf("order %d mode %d tight %d: nodes/q %.1f (p50 %d p99 %d max %d)  leaf tris/q %.1f  exact/q %.1f  mismatches %lld\n", order, mode, tight, (double)nodes / Q,
           pernode[(size_t)Q / 2], pernode[(size_t)(Q * 0.99)], pernode[(size_t)Q - 1], (double)scr / Q, (double)exact / Q, mism);
  }
  return 0;
}
