// clip2d.cuh -- allocation-free integer polygon intersection sweep for the 2D NMS pair test.
//
// What the reference does for every candidate pair (stardist/lib/stardist2d.cpp:152-165,579):
//   Clipper c; c.AddPath(a, ptClip); c.AddPath(b, ptSubject);
//   c.Execute(ctIntersection, res, pftNonZero, pftNonZero);  area = sum_r area_from_path(res[r])
// with the vendored Clipper 6.4.2 (stardist/lib/external/clipper/clipper.cpp).  Clipper is a
// Vatti scanbeam sweep on int64 coordinates that *rounds every edge crossing to the integer
// lattice* (clipper.cpp:136-140, 622-690), so the pair decision `overlap > thresh` depends on
// its exact event order, not only on exact geometry.  Star polygons with r ~ 10 px have ~2 px
// edges on an integer lattice: horizontal edges, shared vertices and coincident edges are the
// common case, not the exception.
//
// This file is a from-scratch sweep with the same *observable semantics* for exactly that use:
// two closed paths, intersection, non-zero fill on both, no open paths, no PolyTree, default
// init options (no reverse / strictly-simple / preserve-collinear), |coord| < 2^30 (the
// reference's "loRange", so its plain int64 slope products are what is reproduced).  It is laid
// out for one GPU thread per pair: fixed-capacity index-linked pools instead of heap nodes and
// std:: containers, a sorted-unique array instead of the scanbeam priority queue, and an
// emulation of libstdc++'s introsort for the two places where the reference's result depends
// on std::sort's (unstable) order (clipper.cpp:1251, 2939).
//
// Compiles as plain C++ too (tests/hostcheck builds it with g++ to compare millions of pairs
// against oracle/_ref/libsdref.so on the CPU box) -- that build is test infrastructure only.
// Device build must use -fmad=false for this TU: the reference's doubles are not contracted.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define SD_HD __host__ __device__
#define SD_HD_BIG __host__ __device__ __noinline__   // large routines: keep one copy per kernel
#else
#define SD_HD
#define SD_HD_BIG
#endif

#ifdef SDC_STATS
#define SDC_HW(k, v) do { if ((v) > hw[k]) hw[k] = (v); } while (0)
#else
#define SDC_HW(k, v) do {} while (0)
#endif

namespace sdclip {

typedef int64_t i64;
typedef int16_t ix;          // pool index, -1 == null
#define SDC_NIL ((ix)-1)
#define SDC_UNASSIGNED (-1)
#define SDC_HORIZONTAL (-1.0E+40)

enum { ptSubject = 0, ptClip = 1 };
enum { esLeft = 1, esRight = 2 };
enum { dRightToLeft = 0, dLeftToRight = 1 };

enum ClipErr {
  CLIP_OK = 0,
  CLIP_FAILED = 1,        // reference would have returned succeeded=false (empty solution)
  CLIP_OVERFLOW = 2       // a fixed-capacity pool overflowed: result invalid, caller must raise
};

struct IPt { int32_t x, y; };

SD_HD inline bool pt_eq(const IPt& a, const IPt& b) { return a.x == b.x && a.y == b.y; }

// clipper.cpp:136-140 (half away from zero, via truncation)
SD_HD inline i64 round_haz(double v) { return (v < 0) ? (i64)(v - 0.5) : (i64)(v + 0.5); }
SD_HD inline i64 iabs64(i64 v) { return v < 0 ? -v : v; }

// clipper.cpp:554-575 with UseFullInt64Range == false
SD_HD inline bool slopes_equal3(IPt p1, IPt p2, IPt p3) {
  return (i64)(p1.y - p2.y) * (i64)(p2.x - p3.x) == (i64)(p1.x - p2.x) * (i64)(p2.y - p3.y);
}
SD_HD inline bool slopes_equal4(IPt p1, IPt p2, IPt p3, IPt p4) {
  return (i64)(p1.y - p2.y) * (i64)(p3.x - p4.x) == (i64)(p1.x - p2.x) * (i64)(p3.y - p4.y);
}
// clipper.cpp:584-588
SD_HD inline double get_dx(IPt p1, IPt p2) {
  return (p1.y == p2.y) ? SDC_HORIZONTAL : (double)((i64)p2.x - p1.x) / (double)((i64)p2.y - p1.y);
}
// clipper.cpp:872-877
SD_HD inline bool horz_segments_overlap(i64 a1, i64 a2, i64 b1, i64 b2) {
  if (a1 > a2) { i64 t = a1; a1 = a2; a2 = t; }
  if (b1 > b2) { i64 t = b1; b1 = b2; b2 = t; }
  return (a1 < b2) && (b1 < a2);
}

struct Edge {
  IPt bot, cur, top;
  double dx;
  int8_t poly, side, wdelta;
  int16_t wcnt, wcnt2;
  int16_t outidx;
  ix next, prev, nextlml, nael, pael, nsel, psel;
};
struct OutPt { int32_t x, y; int16_t idx; ix next, prev; };
struct OutRec { int16_t idx; int8_t ishole; ix firstleft, pts, bottom; };
struct Join { ix op1, op2; IPt off; };
struct LocMin { int32_t y; ix left, right; };
struct INode { ix e1, e2; IPt pt; };

// ---------------------------------------------------------------------------------------
// libstdc++ std::sort (introsort + final insertion sort) on an index permutation, comparator
// "key[a] > key[b]" (descending by Y: LocMinSorter clipper.cpp:123-129, IntersectListSort :2921).
// Equal keys: result order is implementation-defined in the standard, but deterministic in
// libstdc++ (the oracle's toolchain); reproduced here so ties order identically.
// ---------------------------------------------------------------------------------------
template <typename T>
struct StdSortDesc {
  const int32_t* key;   // key per element id
  T* a;                 // array of element ids being sorted
  SD_HD bool comp(T x, T y) const { return key[y] < key[x]; }   // "x before y"
  SD_HD void swp(int i, int j) { T t = a[i]; a[i] = a[j]; a[j] = t; }
  SD_HD void unguarded_linear_insert(int last) {
    T val = a[last]; int nxt = last - 1;
    while (comp(val, a[nxt])) { a[last] = a[nxt]; last = nxt; --nxt; }
    a[last] = val;
  }
  SD_HD void insertion_sort(int first, int last) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
      if (comp(a[i], a[first])) {
        T val = a[i];
        for (int k = i; k > first; --k) a[k] = a[k - 1];
        a[first] = val;
      } else unguarded_linear_insert(i);
    }
  }
  SD_HD void move_median_to_first(int r, int x, int y, int z) {
    if (comp(a[x], a[y])) {
      if (comp(a[y], a[z])) swp(r, y); else if (comp(a[x], a[z])) swp(r, z); else swp(r, x);
    } else if (comp(a[x], a[z])) swp(r, x);
    else if (comp(a[y], a[z])) swp(r, z);
    else swp(r, y);
  }
  SD_HD int unguarded_partition(int first, int last, int pivot) {
    for (;;) {
      while (comp(a[first], a[pivot])) ++first;
      --last;
      while (comp(a[pivot], a[last])) --last;
      if (!(first < last)) return first;
      swp(first, last);
      ++first;
    }
  }
  // heap helpers for the depth-limit fallback (std::__partial_sort(first,last,last))
  SD_HD void push_heap_(int first, int hole, int top, T val) {
    int parent = (hole - 1) / 2;
    while (hole > top && comp(a[first + parent], val)) {
      a[first + hole] = a[first + parent]; hole = parent; parent = (hole - 1) / 2;
    }
    a[first + hole] = val;
  }
  SD_HD void adjust_heap(int first, int hole, int len, T val) {
    const int top = hole; int child = hole;
    while (child < (len - 1) / 2) {
      child = 2 * (child + 1);
      if (comp(a[first + child], a[first + child - 1])) child--;
      a[first + hole] = a[first + child]; hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
      child = 2 * (child + 1);
      a[first + hole] = a[first + child - 1]; hole = child - 1;
    }
    push_heap_(first, hole, top, val);
  }
  SD_HD void heap_sort(int first, int last) {
    int len = last - first;
    if (len >= 2) {
      int parent = (len - 2) / 2;
      for (;;) { T v = a[first + parent]; adjust_heap(first, parent, len, v); if (parent == 0) break; parent--; }
    }
    while (last - first > 1) {
      --last;
      T v = a[last]; a[last] = a[first];
      adjust_heap(first, 0, last - first, v);
    }
  }
  SD_HD_BIG void sort(int n) {
    if (n <= 1) return;
    // __introsort_loop, recursion unrolled with an explicit stack of (first,last,depth)
    int stf[40], stl[40], std_[40]; int sp = 0;
    int lg = 0; for (int t = n; t > 1; t >>= 1) lg++;
    stf[0] = 0; stl[0] = n; std_[0] = 2 * lg; sp = 1;
    while (sp > 0) {
      --sp; int first = stf[sp], last = stl[sp], depth = std_[sp];
      while (last - first > 16) {
        if (depth == 0) { heap_sort(first, last); break; }
        --depth;
        int mid = first + (last - first) / 2;
        move_median_to_first(first, first + 1, mid, last - 1);
        int cut = unguarded_partition(first + 1, last, first);
        // recurse on [cut,last) first (pushed; order of processing does not affect the result
        // because the ranges are disjoint), loop on [first,cut)
        if (sp < 40) { stf[sp] = cut; stl[sp] = last; std_[sp] = depth; sp++; }
        last = cut;
      }
    }
    if (n > 16) {
      insertion_sort(0, 16);
      for (int i = 16; i != n; ++i) unguarded_linear_insert(i);
    } else insertion_sort(0, n);
  }
};

// ---------------------------------------------------------------------------------------
// NV = max vertices per polygon; SC scales the pools whose size is data dependent.  The GPU
// fast path runs SC=1 (high-water marks over 1.5M fuzzed 32-gon pairs: P 84, R 22, J 6, GJ 6,
// IL 17); a pair that overflows is re-run by the slow path with SC=4 -- still on the device.
template <int NV, int SC = 4>
struct ClipSweep {
  enum {
    MAXE  = 2 * NV,
    MAXOP = 4 * NV * SC,
    MAXOR = NV * SC,
    MAXJ  = (NV / 2) * SC + 8,
    MAXGJ = (NV / 2) * SC + 8,
    MAXLM = NV + 2,            // hard bound: a closed n-gon has at most n/2 local minima
    MAXSB = 2 * NV + 4,        // hard bound: distinct vertex Ys
    MAXIL = 2 * NV * SC
  };
  Edge E[MAXE];
  OutPt P[MAXOP];
  OutRec R[MAXOR];
  Join J[MAXJ];
  Join GJ[MAXGJ];
  LocMin LM[MAXLM];
  int32_t SB[MAXSB];      // scanbeam: sorted ascending, unique (== priority_queue + dup-pop)
  INode IL[MAXIL];
  int nE, nP, nR, nJ, nGJ, nLM, curLM, nSB, nIL;
  ix ael, sel;
  int err;
#ifdef SDC_STATS
  int hw[8];   // high-water marks: P,R,J,GJ,LM,SB,IL,E
#endif

  SD_HD void init() {
    nE = nP = nR = nJ = nGJ = nLM = curLM = nSB = nIL = 0;
    ael = sel = SDC_NIL; err = CLIP_OK;
#ifdef SDC_STATS
    for (int i = 0; i < 8; ++i) hw[i] = 0;
#endif
  }
  SD_HD void fail(int e) { if (err < e) err = e; }

  // ------------------------------------------------------------------ small helpers
  SD_HD bool is_horz(ix e) const { return E[e].dx == SDC_HORIZONTAL; }
  SD_HD void set_dx(ix e) {                                      // clipper.cpp:591-596
    i64 dy = (i64)E[e].top.y - E[e].bot.y;
    if (dy == 0) E[e].dx = SDC_HORIZONTAL;
    else E[e].dx = (double)((i64)E[e].top.x - E[e].bot.x) / (double)dy;
  }
  SD_HD i64 top_x(ix e, i64 y) const {                            // clipper.cpp:615-619
    const Edge& ed = E[e];
    return (y == ed.top.y) ? (i64)ed.top.x : (i64)ed.bot.x + round_haz(ed.dx * (double)(y - ed.bot.y));
  }
  SD_HD void reverse_horizontal(ix e) { int32_t t = E[e].top.x; E[e].top.x = E[e].bot.x; E[e].bot.x = t; }
  SD_HD IPt opt(ix p) const { IPt r; r.x = P[p].x; r.y = P[p].y; return r; }

  SD_HD void sb_insert(i64 y64) {
    int32_t y = (int32_t)y64;
    int lo = 0, hi = nSB;
    while (lo < hi) { int m = (lo + hi) >> 1; if (SB[m] < y) lo = m + 1; else hi = m; }
    if (lo < nSB && SB[lo] == y) return;
    if (nSB >= MAXSB) { fail(CLIP_OVERFLOW); return; }
    for (int k = nSB; k > lo; --k) SB[k] = SB[k - 1];
    SB[lo] = y; nSB++; SDC_HW(5, nSB);
  }
  SD_HD bool sb_pop(i64& y) { if (nSB == 0) return false; y = SB[--nSB]; return true; }

  // ------------------------------------------------------------------ AddPath (clipper.cpp:1045-1221)
  SD_HD ix remove_edge(ix e) {
    E[E[e].prev].next = E[e].next;
    E[E[e].next].prev = E[e].prev;
    ix r = E[e].next;
    E[e].prev = SDC_NIL;
    return r;
  }
  SD_HD ix find_next_loc_min(ix e) {                              // clipper.cpp:911-925
    for (;;) {
      while (!pt_eq(E[e].bot, E[E[e].prev].bot) || pt_eq(E[e].cur, E[e].top)) e = E[e].next;
      if (!is_horz(e) && !is_horz(E[e].prev)) break;
      while (is_horz(E[e].prev)) e = E[e].prev;
      ix e2 = e;
      while (is_horz(e)) e = E[e].next;
      if (E[e].top.y == E[E[e].prev].bot.y) continue;
      if (E[E[e2].prev].bot.x < E[e].bot.x) e = e2;
      break;
    }
    return e;
  }
  SD_HD ix process_bound(ix e, bool fwd) {                        // clipper.cpp:928-1042 (closed paths)
    ix result = e, horz;
    if (is_horz(e)) {
      ix es = fwd ? E[e].prev : E[e].next;
      if (is_horz(es)) {
        if (E[es].bot.x != E[e].bot.x && E[es].top.x != E[e].bot.x) reverse_horizontal(e);
      } else if (E[es].bot.x != E[e].bot.x) reverse_horizontal(e);
    }
    ix estart = e;
    if (fwd) {
      while (E[result].top.y == E[E[result].next].bot.y) result = E[result].next;
      if (is_horz(result)) {
        horz = result;
        while (is_horz(E[horz].prev)) horz = E[horz].prev;
        if (E[E[horz].prev].top.x > E[E[result].next].top.x) result = E[horz].prev;
      }
      while (e != result) {
        E[e].nextlml = E[e].next;
        if (is_horz(e) && e != estart && E[e].bot.x != E[E[e].prev].top.x) reverse_horizontal(e);
        e = E[e].next;
      }
      if (is_horz(e) && e != estart && E[e].bot.x != E[E[e].prev].top.x) reverse_horizontal(e);
      result = E[result].next;
    } else {
      while (E[result].top.y == E[E[result].prev].bot.y) result = E[result].prev;
      if (is_horz(result)) {
        horz = result;
        while (is_horz(E[horz].next)) horz = E[horz].next;
        if (E[E[horz].next].top.x == E[E[result].prev].top.x ||
            E[E[horz].next].top.x > E[E[result].prev].top.x) result = E[horz].next;
      }
      while (e != result) {
        E[e].nextlml = E[e].prev;
        if (is_horz(e) && e != estart && E[e].bot.x != E[E[e].next].top.x) reverse_horizontal(e);
        e = E[e].prev;
      }
      if (is_horz(e) && e != estart && E[e].bot.x != E[E[e].next].top.x) reverse_horizontal(e);
      result = E[result].prev;
    }
    return result;
  }

  // v: vertex source with v.x(i), v.y(i) for i in [0,n) (already truncated to integers by the
  // caller, stardist2d.cpp:471)
  template <typename VSrc>
  SD_HD_BIG bool add_path(const VSrc& v, int n, int poly) {
    int hi = n - 1;
    while (hi > 0 && v.x(hi) == v.x(0) && v.y(hi) == v.y(0)) --hi;
    while (hi > 0 && v.x(hi) == v.x(hi - 1) && v.y(hi) == v.y(hi - 1)) --hi;
    if (hi < 2) return false;
    if (nE + hi + 1 > MAXE) { fail(CLIP_OVERFLOW); return false; }
    const int base = nE;
    for (int i = 0; i <= hi; ++i) {
      Edge& e = E[base + i];
      e.bot.x = e.bot.y = e.top.x = e.top.y = 0; e.dx = 0.0;
      e.cur.x = (int32_t)v.x(i); e.cur.y = (int32_t)v.y(i);
      e.poly = 0; e.side = 0; e.wdelta = 0; e.wcnt = 0; e.wcnt2 = 0;
      e.outidx = SDC_UNASSIGNED;
      e.next = (ix)(base + (i == hi ? 0 : i + 1));
      e.prev = (ix)(base + (i == 0 ? hi : i - 1));
      e.nextlml = e.nael = e.pael = e.nsel = e.psel = SDC_NIL;
    }
    nE += hi + 1; SDC_HW(7, nE);      // slots stay reserved even if the path is rejected below (like m_edges)
    ix estart = (ix)base, e = estart, eloopstop = estart;
    for (;;) {
      if (pt_eq(E[e].cur, E[E[e].next].cur)) {
        if (e == E[e].next) break;
        if (e == estart) estart = E[e].next;
        e = remove_edge(e);
        eloopstop = e;
        continue;
      }
      if (E[e].prev == E[e].next) break;
      else if (slopes_equal3(E[E[e].prev].cur, E[e].cur, E[E[e].next].cur)) {
        if (e == estart) estart = E[e].next;
        e = remove_edge(e);
        e = E[e].prev;
        eloopstop = e;
        continue;
      }
      e = E[e].next;
      if (e == eloopstop) break;
    }
    if (E[e].prev == E[e].next) { nE = base; return false; }

    bool isflat = true;
    e = estart;
    do {
      // InitEdge2 (clipper.cpp:729-742)
      ix nx = E[e].next;
      if (E[e].cur.y >= E[nx].cur.y) { E[e].bot = E[e].cur; E[e].top = E[nx].cur; }
      else { E[e].top = E[e].cur; E[e].bot = E[nx].cur; }
      set_dx(e);
      E[e].poly = (int8_t)poly;
      e = E[e].next;
      if (isflat && E[e].cur.y != E[estart].cur.y) isflat = false;
    } while (e != estart);
    if (isflat) { nE = base; return false; }

    bool lfwd;
    ix emin = SDC_NIL;
    if (pt_eq(E[E[e].prev].bot, E[E[e].prev].top)) e = E[e].next;
    for (;;) {
      e = find_next_loc_min(e);
      if (e == emin) break;
      else if (emin == SDC_NIL) emin = e;
      if (nLM >= MAXLM) { fail(CLIP_OVERFLOW); return false; }
      LocMin& lm = LM[nLM];
      lm.y = E[e].bot.y;
      if (E[e].dx < E[E[e].prev].dx) { lm.left = E[e].prev; lm.right = e; lfwd = false; }
      else { lm.left = e; lm.right = E[e].prev; lfwd = true; }
      if (E[lm.left].next == lm.right) E[lm.left].wdelta = -1; else E[lm.left].wdelta = 1;
      E[lm.right].wdelta = (int8_t)(-E[lm.left].wdelta);
      e = process_bound(lm.left, lfwd);
      ix e2 = process_bound(lm.right, !lfwd);
      nLM++; SDC_HW(4, nLM);
      if (!lfwd) e = e2;
    }
    return true;
  }

  // ------------------------------------------------------------------ Reset (clipper.cpp:1247-1276)
  SD_HD_BIG void reset() {
    curLM = 0;
    if (nLM == 0) return;
    {
      int32_t keys[MAXLM]; int16_t perm[MAXLM]; LocMin tmp[MAXLM];
      for (int i = 0; i < nLM; ++i) { keys[i] = LM[i].y; perm[i] = (int16_t)i; tmp[i] = LM[i]; }
      StdSortDesc<int16_t> s; s.key = keys; s.a = perm; s.sort(nLM);
      for (int i = 0; i < nLM; ++i) LM[i] = tmp[perm[i]];
    }
    nSB = 0;
    for (int i = 0; i < nLM; ++i) {
      sb_insert(LM[i].y);
      ix e = LM[i].left;
      E[e].cur = E[e].bot; E[e].side = esLeft; E[e].outidx = SDC_UNASSIGNED;
      e = LM[i].right;
      E[e].cur = E[e].bot; E[e].side = esRight; E[e].outidx = SDC_UNASSIGNED;
    }
    ael = SDC_NIL;
    curLM = 0;
  }

  // ------------------------------------------------------------------ AEL / SEL plumbing
  SD_HD void delete_from_ael(ix e) {                              // clipper.cpp:1367-1377
    ix p = E[e].pael, n = E[e].nael;
    if (p == SDC_NIL && n == SDC_NIL && e != ael) return;
    if (p != SDC_NIL) E[p].nael = n; else ael = n;
    if (n != SDC_NIL) E[n].pael = p;
    E[e].nael = SDC_NIL; E[e].pael = SDC_NIL;
  }
  SD_HD void delete_from_sel(ix e) {                              // clipper.cpp:2080-2090
    ix p = E[e].psel, n = E[e].nsel;
    if (p == SDC_NIL && n == SDC_NIL && e != sel) return;
    if (p != SDC_NIL) E[p].nsel = n; else sel = n;
    if (n != SDC_NIL) E[n].psel = p;
    E[e].nsel = SDC_NIL; E[e].psel = SDC_NIL;
  }
  SD_HD_BIG void swap_positions_in_ael(ix e1, ix e2) {                // clipper.cpp:1395-1439
    if (E[e1].nael == E[e1].pael || E[e2].nael == E[e2].pael) return;
    if (E[e1].nael == e2) {
      ix n = E[e2].nael; if (n != SDC_NIL) E[n].pael = e1;
      ix p = E[e1].pael; if (p != SDC_NIL) E[p].nael = e2;
      E[e2].pael = p; E[e2].nael = e1; E[e1].pael = e2; E[e1].nael = n;
    } else if (E[e2].nael == e1) {
      ix n = E[e1].nael; if (n != SDC_NIL) E[n].pael = e2;
      ix p = E[e2].pael; if (p != SDC_NIL) E[p].nael = e1;
      E[e1].pael = p; E[e1].nael = e2; E[e2].pael = e1; E[e2].nael = n;
    } else {
      ix n = E[e1].nael, p = E[e1].pael;
      E[e1].nael = E[e2].nael; if (E[e1].nael != SDC_NIL) E[E[e1].nael].pael = e1;
      E[e1].pael = E[e2].pael; if (E[e1].pael != SDC_NIL) E[E[e1].pael].nael = e1;
      E[e2].nael = n; if (n != SDC_NIL) E[n].pael = e2;
      E[e2].pael = p; if (p != SDC_NIL) E[p].nael = e2;
    }
    if (E[e1].pael == SDC_NIL) ael = e1;
    else if (E[e2].pael == SDC_NIL) ael = e2;
  }
  SD_HD_BIG void swap_positions_in_sel(ix e1, ix e2) {                // clipper.cpp:2558-2601
    if (E[e1].nsel == SDC_NIL && E[e1].psel == SDC_NIL) return;
    if (E[e2].nsel == SDC_NIL && E[e2].psel == SDC_NIL) return;
    if (E[e1].nsel == e2) {
      ix n = E[e2].nsel; if (n != SDC_NIL) E[n].psel = e1;
      ix p = E[e1].psel; if (p != SDC_NIL) E[p].nsel = e2;
      E[e2].psel = p; E[e2].nsel = e1; E[e1].psel = e2; E[e1].nsel = n;
    } else if (E[e2].nsel == e1) {
      ix n = E[e1].nsel; if (n != SDC_NIL) E[n].psel = e2;
      ix p = E[e2].psel; if (p != SDC_NIL) E[p].nsel = e1;
      E[e1].psel = p; E[e1].nsel = e2; E[e2].psel = e1; E[e2].nsel = n;
    } else {
      ix n = E[e1].nsel, p = E[e1].psel;
      E[e1].nsel = E[e2].nsel; if (E[e1].nsel != SDC_NIL) E[E[e1].nsel].psel = e1;
      E[e1].psel = E[e2].psel; if (E[e1].psel != SDC_NIL) E[E[e1].psel].nsel = e1;
      E[e2].nsel = n; if (n != SDC_NIL) E[n].psel = e2;
      E[e2].psel = p; if (p != SDC_NIL) E[p].nsel = e2;
    }
    if (E[e1].psel == SDC_NIL) sel = e1;
    else if (E[e2].psel == SDC_NIL) sel = e2;
  }
  SD_HD bool update_edge_into_ael(ix& e) {                        // clipper.cpp:1442-1462
    ix nl = E[e].nextlml;
    if (nl == SDC_NIL) { fail(CLIP_FAILED); return false; }
    E[nl].outidx = E[e].outidx;
    ix p = E[e].pael, n = E[e].nael;
    if (p != SDC_NIL) E[p].nael = nl; else ael = nl;
    if (n != SDC_NIL) E[n].pael = nl;
    E[nl].side = E[e].side; E[nl].wdelta = E[e].wdelta;
    E[nl].wcnt = E[e].wcnt; E[nl].wcnt2 = E[e].wcnt2;
    e = nl;
    E[e].cur = E[e].bot;
    E[e].pael = p; E[e].nael = n;
    if (!is_horz(e)) sb_insert(E[e].top.y);
    return true;
  }
  SD_HD bool e2_inserts_before_e1(ix e1, ix e2) const {           // clipper.cpp:3278-3287
    if (E[e2].cur.x == E[e1].cur.x) {
      if (E[e2].top.y > E[e1].top.y) return (i64)E[e2].top.x < top_x(e1, E[e2].top.y);
      else return (i64)E[e1].top.x > top_x(e2, E[e1].top.y);
    } else return E[e2].cur.x < E[e1].cur.x;
  }
  SD_HD void insert_edge_into_ael(ix edge, ix start) {            // clipper.cpp:3319-3345
    if (ael == SDC_NIL) {
      E[edge].pael = SDC_NIL; E[edge].nael = SDC_NIL; ael = edge;
    } else if (start == SDC_NIL && e2_inserts_before_e1(ael, edge)) {
      E[edge].pael = SDC_NIL; E[edge].nael = ael; E[ael].pael = edge; ael = edge;
    } else {
      if (start == SDC_NIL) start = ael;
      while (E[start].nael != SDC_NIL && !e2_inserts_before_e1(E[start].nael, edge)) start = E[start].nael;
      E[edge].nael = E[start].nael;
      if (E[start].nael != SDC_NIL) E[E[start].nael].pael = edge;
      E[edge].pael = start;
      E[start].nael = edge;
    }
  }
  SD_HD void add_edge_to_sel(ix e) {                              // clipper.cpp:1900-1917
    if (sel == SDC_NIL) { sel = e; E[e].psel = SDC_NIL; E[e].nsel = SDC_NIL; }
    else { E[e].nsel = sel; E[e].psel = SDC_NIL; E[sel].psel = e; sel = e; }
  }

  // ------------------------------------------------------------------ output records
  SD_HD ix new_outpt() {
    if (nP >= MAXOP) { fail(CLIP_OVERFLOW); return (ix)(MAXOP - 1); }
    nP++; SDC_HW(0, nP);
    return (ix)(nP - 1);
  }
  SD_HD ix create_outrec() {                                      // clipper.cpp:1380-1392
    if (nR >= MAXOR) { fail(CLIP_OVERFLOW); return (ix)(MAXOR - 1); }
    OutRec& r = R[nR];
    r.ishole = 0; r.firstleft = SDC_NIL; r.pts = SDC_NIL; r.bottom = SDC_NIL; r.idx = (int16_t)nR;
    nR++; SDC_HW(1, nR);
    return (ix)(nR - 1);
  }
  SD_HD void set_hole_state(ix e, ix orec) {                      // clipper.cpp:2301-2324
    ix e2 = E[e].pael, tmp = SDC_NIL;
    while (e2 != SDC_NIL) {
      if (E[e2].outidx >= 0) {
        if (tmp == SDC_NIL) tmp = e2;
        else if (E[tmp].outidx == E[e2].outidx) tmp = SDC_NIL;
      }
      e2 = E[e2].pael;
    }
    if (tmp == SDC_NIL) { R[orec].firstleft = SDC_NIL; R[orec].ishole = 0; }
    else { R[orec].firstleft = (ix)E[tmp].outidx; R[orec].ishole = (int8_t)!R[R[orec].firstleft].ishole; }
  }
  SD_HD_BIG ix add_outpt(ix e, IPt pt) {                              // clipper.cpp:2463-2499
    if (E[e].outidx < 0) {
      ix orec = create_outrec();
      ix np = new_outpt();
      R[orec].pts = np;
      P[np].idx = R[orec].idx; P[np].x = pt.x; P[np].y = pt.y; P[np].next = np; P[np].prev = np;
      set_hole_state(e, orec);
      E[e].outidx = R[orec].idx;
      return np;
    } else {
      ix orec = (ix)E[e].outidx;
      ix op = R[orec].pts;
      bool tofront = (E[e].side == esLeft);
      if (tofront && pt.x == P[op].x && pt.y == P[op].y) return op;
      else if (!tofront && pt.x == P[P[op].prev].x && pt.y == P[P[op].prev].y) return P[op].prev;
      ix np = new_outpt();
      P[np].idx = R[orec].idx; P[np].x = pt.x; P[np].y = pt.y;
      P[np].next = op; P[np].prev = P[op].prev;
      P[P[np].prev].next = np;
      P[op].prev = np;
      if (tofront) R[orec].pts = np;
      return np;
    }
  }
  SD_HD ix get_last_outpt(ix e) const {                           // clipper.cpp:2502-2509
    ix orec = (ix)E[e].outidx;
    if (E[e].side == esLeft) return R[orec].pts; else return P[R[orec].pts].prev;
  }
  SD_HD void add_join(ix op1, ix op2, IPt off) {
    if (nJ >= MAXJ) { fail(CLIP_OVERFLOW); return; }
    J[nJ].op1 = op1; J[nJ].op2 = op2; J[nJ].off = off; nJ++; SDC_HW(2, nJ);
  }
  SD_HD void add_ghost_join(ix op, IPt off) {
    if (nGJ >= MAXGJ) { fail(CLIP_OVERFLOW); return; }
    GJ[nGJ].op1 = op; GJ[nGJ].op2 = SDC_NIL; GJ[nGJ].off = off; nGJ++; SDC_HW(3, nGJ);
  }
  SD_HD void reverse_poly_pt_links(ix pp) {                       // clipper.cpp:692-703
    if (pp == SDC_NIL) return;
    ix p1 = pp;
    do { ix p2 = P[p1].next; P[p1].next = P[p1].prev; P[p1].prev = p2; p1 = p2; } while (p1 != pp);
  }
  SD_HD double area_op(ix op) const {                             // clipper.cpp:406-416
    ix start = op;
    if (op == SDC_NIL) return 0;
    double a = 0;
    do {
      a += (double)((i64)P[P[op].prev].x + P[op].x) * (double)((i64)P[P[op].prev].y - P[op].y);
      op = P[op].next;
    } while (op != start);
    return a * 0.5;
  }
  SD_HD_BIG bool first_is_bottom_pt(ix b1, ix b2) const {             // clipper.cpp:798-819
    ix p = P[b1].prev;
    while (pt_eq(opt(p), opt(b1)) && p != b1) p = P[p].prev;
    double dx1p = fabs(get_dx(opt(b1), opt(p)));
    p = P[b1].next;
    while (pt_eq(opt(p), opt(b1)) && p != b1) p = P[p].next;
    double dx1n = fabs(get_dx(opt(b1), opt(p)));
    p = P[b2].prev;
    while (pt_eq(opt(p), opt(b2)) && p != b2) p = P[p].prev;
    double dx2p = fabs(get_dx(opt(b2), opt(p)));
    p = P[b2].next;
    while (pt_eq(opt(p), opt(b2)) && p != b2) p = P[p].next;
    double dx2n = fabs(get_dx(opt(b2), opt(p)));
    double mx1 = dx1p < dx1n ? dx1n : dx1p, mn1 = dx1n < dx1p ? dx1n : dx1p;   // std::max/min
    double mx2 = dx2p < dx2n ? dx2n : dx2p, mn2 = dx2n < dx2p ? dx2n : dx2p;
    if (mx1 == mx2 && mn1 == mn2) return area_op(b1) > 0;
    else return (dx1p >= dx2p && dx1p >= dx2n) || (dx1n >= dx2p && dx1n >= dx2n);
  }
  SD_HD_BIG ix get_bottom_pt(ix pp) const {                           // clipper.cpp:822-857
    ix dups = SDC_NIL;
    ix p = P[pp].next;
    while (p != pp) {
      if (P[p].y > P[pp].y) { pp = p; dups = SDC_NIL; }
      else if (P[p].y == P[pp].y && P[p].x <= P[pp].x) {
        if (P[p].x < P[pp].x) { dups = SDC_NIL; pp = p; }
        else { if (P[p].next != pp && P[p].prev != pp) dups = p; }
      }
      p = P[p].next;
    }
    if (dups != SDC_NIL) {
      while (dups != p) {
        if (!first_is_bottom_pt(p, dups)) pp = dups;
        dups = P[dups].next;
        while (!pt_eq(opt(dups), opt(pp))) dups = P[dups].next;
      }
    }
    return pp;
  }
  SD_HD ix get_lowermost_rec(ix r1, ix r2) {                      // clipper.cpp:2327-2344
    if (R[r1].bottom == SDC_NIL) R[r1].bottom = get_bottom_pt(R[r1].pts);
    if (R[r2].bottom == SDC_NIL) R[r2].bottom = get_bottom_pt(R[r2].pts);
    ix o1 = R[r1].bottom, o2 = R[r2].bottom;
    if (P[o1].y > P[o2].y) return r1;
    else if (P[o1].y < P[o2].y) return r2;
    else if (P[o1].x < P[o2].x) return r1;
    else if (P[o1].x > P[o2].x) return r2;
    else if (P[o1].next == o1) return r2;
    else if (P[o2].next == o2) return r1;
    else if (first_is_bottom_pt(o1, o2)) return r1;
    else return r2;
  }
  SD_HD bool outrec1_right_of_outrec2(ix r1, ix r2) const {       // clipper.cpp:2347-2355
    do { r1 = R[r1].firstleft; if (r1 == r2) return true; } while (r1 != SDC_NIL);
    return false;
  }
  SD_HD ix get_outrec(int idx) const {                            // clipper.cpp:2358-2364
    ix r = (ix)idx;
    while (r != R[r].idx) r = (ix)R[r].idx;
    return r;
  }
  SD_HD_BIG void append_polygon(ix e1, ix e2) {                       // clipper.cpp:2367-2460
    ix r1 = (ix)E[e1].outidx, r2 = (ix)E[e2].outidx;
    ix holerec;
    if (outrec1_right_of_outrec2(r1, r2)) holerec = r2;
    else if (outrec1_right_of_outrec2(r2, r1)) holerec = r1;
    else holerec = get_lowermost_rec(r1, r2);
    ix p1l = R[r1].pts, p1r = P[p1l].prev, p2l = R[r2].pts, p2r = P[p2l].prev;
    if (E[e1].side == esLeft) {
      if (E[e2].side == esLeft) {
        reverse_poly_pt_links(p2l);
        P[p2l].next = p1l; P[p1l].prev = p2l; P[p1r].next = p2r; P[p2r].prev = p1r;
        R[r1].pts = p2r;
      } else {
        P[p2r].next = p1l; P[p1l].prev = p2r; P[p2l].prev = p1r; P[p1r].next = p2l;
        R[r1].pts = p2l;
      }
    } else {
      if (E[e2].side == esRight) {
        reverse_poly_pt_links(p2l);
        P[p1r].next = p2r; P[p2r].prev = p1r; P[p2l].next = p1l; P[p1l].prev = p2l;
      } else {
        P[p1r].next = p2l; P[p2l].prev = p1r; P[p1l].prev = p2r; P[p2r].next = p1l;
      }
    }
    R[r1].bottom = SDC_NIL;
    if (holerec == r2) {
      if (R[r2].firstleft != r1) R[r1].firstleft = R[r2].firstleft;
      R[r1].ishole = R[r2].ishole;
    }
    R[r2].pts = SDC_NIL; R[r2].bottom = SDC_NIL; R[r2].firstleft = r1;
    int ok = E[e1].outidx, obsolete = E[e2].outidx;
    E[e1].outidx = SDC_UNASSIGNED; E[e2].outidx = SDC_UNASSIGNED;
    ix e = ael;
    while (e != SDC_NIL) {
      if (E[e].outidx == obsolete) { E[e].outidx = (int16_t)ok; E[e].side = E[e1].side; break; }
      e = E[e].nael;
    }
    R[r2].idx = R[r1].idx;
  }
  SD_HD_BIG ix add_local_min_poly(ix e1, ix e2, IPt pt) {             // clipper.cpp:1841-1881
    ix result, e, preve;
    if (is_horz(e2) || (E[e1].dx > E[e2].dx)) {
      result = add_outpt(e1, pt);
      E[e2].outidx = E[e1].outidx; E[e1].side = esLeft; E[e2].side = esRight;
      e = e1;
      preve = (E[e].pael == e2) ? E[e2].pael : E[e].pael;
    } else {
      result = add_outpt(e2, pt);
      E[e1].outidx = E[e2].outidx; E[e1].side = esRight; E[e2].side = esLeft;
      e = e2;
      preve = (E[e].pael == e1) ? E[e1].pael : E[e].pael;
    }
    if (preve != SDC_NIL && E[preve].outidx >= 0 && E[preve].top.y < pt.y && E[e].top.y < pt.y) {
      i64 xp = top_x(preve, pt.y), xe = top_x(e, pt.y);
      IPt a; a.x = (int32_t)xp; a.y = pt.y; IPt b; b.x = (int32_t)xe; b.y = pt.y;
      if (xp == xe && slopes_equal4(a, E[preve].top, b, E[e].top)) {
        ix op = add_outpt(preve, pt);
        add_join(result, op, E[e].top);
      }
    }
    return result;
  }
  SD_HD void add_local_max_poly(ix e1, ix e2, IPt pt) {           // clipper.cpp:1884-1897
    add_outpt(e1, pt);
    if (E[e1].outidx == E[e2].outidx) { E[e1].outidx = SDC_UNASSIGNED; E[e2].outidx = SDC_UNASSIGNED; }
    else if (E[e1].outidx < E[e2].outidx) append_polygon(e1, e2);
    else append_polygon(e2, e1);
  }

  // ------------------------------------------------------------------ winding (non-zero only)
  SD_HD void set_winding_count(ix edge) {                         // clipper.cpp:1624-1722
    ix e = E[edge].pael;
    while (e != SDC_NIL && E[e].poly != E[edge].poly) e = E[e].pael;
    if (e == SDC_NIL) {
      E[edge].wcnt = E[edge].wdelta; E[edge].wcnt2 = 0; e = ael;
    } else {
      if ((int)E[e].wcnt * E[e].wdelta < 0) {
        int aw = E[e].wcnt < 0 ? -E[e].wcnt : E[e].wcnt;
        if (aw > 1) {
          if (E[e].wdelta * E[edge].wdelta < 0) E[edge].wcnt = E[e].wcnt;
          else E[edge].wcnt = (int16_t)(E[e].wcnt + E[edge].wdelta);
        } else E[edge].wcnt = E[edge].wdelta;
      } else {
        if (E[e].wdelta * E[edge].wdelta < 0) E[edge].wcnt = E[e].wcnt;
        else E[edge].wcnt = (int16_t)(E[e].wcnt + E[edge].wdelta);
      }
      E[edge].wcnt2 = E[e].wcnt2;
      e = E[e].nael;
    }
    while (e != edge) { E[edge].wcnt2 = (int16_t)(E[edge].wcnt2 + E[e].wdelta); e = E[e].nael; }
  }
  SD_HD bool is_contributing(ix e) const {                        // clipper.cpp:1741-1838
    int aw = E[e].wcnt < 0 ? -E[e].wcnt : E[e].wcnt;
    if (aw != 1) return false;
    return E[e].wcnt2 != 0;
  }
  SD_HD_BIG void intersect_edges(ix e1, ix e2, IPt pt) {              // clipper.cpp:2106-2298
    bool c1 = E[e1].outidx >= 0, c2 = E[e2].outidx >= 0;
    if (E[e1].poly == E[e2].poly) {
      if (E[e1].wcnt + E[e2].wdelta == 0) E[e1].wcnt = (int16_t)(-E[e1].wcnt);
      else E[e1].wcnt = (int16_t)(E[e1].wcnt + E[e2].wdelta);
      if (E[e2].wcnt - E[e1].wdelta == 0) E[e2].wcnt = (int16_t)(-E[e2].wcnt);
      else E[e2].wcnt = (int16_t)(E[e2].wcnt - E[e1].wdelta);
    } else {
      E[e1].wcnt2 = (int16_t)(E[e1].wcnt2 + E[e2].wdelta);
      E[e2].wcnt2 = (int16_t)(E[e2].wcnt2 - E[e1].wdelta);
    }
    int w1 = E[e1].wcnt < 0 ? -E[e1].wcnt : E[e1].wcnt;
    int w2 = E[e2].wcnt < 0 ? -E[e2].wcnt : E[e2].wcnt;
    if (c1 && c2) {
      if ((w1 != 0 && w1 != 1) || (w2 != 0 && w2 != 1) || (E[e1].poly != E[e2].poly)) {
        add_local_max_poly(e1, e2, pt);
      } else {
        add_outpt(e1, pt); add_outpt(e2, pt);
        int8_t s = E[e1].side; E[e1].side = E[e2].side; E[e2].side = s;
        int16_t o = E[e1].outidx; E[e1].outidx = E[e2].outidx; E[e2].outidx = o;
      }
    } else if (c1) {
      if (w2 == 0 || w2 == 1) {
        add_outpt(e1, pt);
        int8_t s = E[e1].side; E[e1].side = E[e2].side; E[e2].side = s;
        int16_t o = E[e1].outidx; E[e1].outidx = E[e2].outidx; E[e2].outidx = o;
      }
    } else if (c2) {
      if (w1 == 0 || w1 == 1) {
        add_outpt(e2, pt);
        int8_t s = E[e1].side; E[e1].side = E[e2].side; E[e2].side = s;
        int16_t o = E[e1].outidx; E[e1].outidx = E[e2].outidx; E[e2].outidx = o;
      }
    } else if ((w1 == 0 || w1 == 1) && (w2 == 0 || w2 == 1)) {
      int w12 = E[e1].wcnt2 < 0 ? -E[e1].wcnt2 : E[e1].wcnt2;
      int w22 = E[e2].wcnt2 < 0 ? -E[e2].wcnt2 : E[e2].wcnt2;
      if (E[e1].poly != E[e2].poly) add_local_min_poly(e1, e2, pt);
      else if (w1 == 1 && w2 == 1) { if (w12 > 0 && w22 > 0) add_local_min_poly(e1, e2, pt); }
      else { int8_t s = E[e1].side; E[e1].side = E[e2].side; E[e2].side = s; }
    }
  }

  // ------------------------------------------------------------------ InsertLocalMinimaIntoAEL (clipper.cpp:1978-2077)
  SD_HD_BIG void insert_local_minima_into_ael(i64 boty) {
    while (curLM < nLM && LM[curLM].y == boty) {
      ix lb = LM[curLM].left, rb = LM[curLM].right;
      curLM++;
      ix op1 = SDC_NIL;
      insert_edge_into_ael(lb, SDC_NIL);
      insert_edge_into_ael(rb, lb);
      set_winding_count(lb);
      E[rb].wcnt = E[lb].wcnt; E[rb].wcnt2 = E[lb].wcnt2;
      if (is_contributing(lb)) op1 = add_local_min_poly(lb, rb, E[lb].bot);
      sb_insert(E[lb].top.y);
      if (is_horz(rb)) {
        add_edge_to_sel(rb);
        if (E[rb].nextlml != SDC_NIL) sb_insert(E[E[rb].nextlml].top.y);
      } else sb_insert(E[rb].top.y);

      if (op1 != SDC_NIL && is_horz(rb) && nGJ > 0) {
        for (int i = 0; i < nGJ; ++i) {
          if (horz_segments_overlap(P[GJ[i].op1].x, GJ[i].off.x, E[rb].bot.x, E[rb].top.x))
            add_join(GJ[i].op1, op1, GJ[i].off);
        }
      }
      ix lp = E[lb].pael;
      if (E[lb].outidx >= 0 && lp != SDC_NIL && E[lp].cur.x == E[lb].bot.x && E[lp].outidx >= 0 &&
          slopes_equal4(E[lp].bot, E[lp].top, E[lb].cur, E[lb].top)) {
        ix op2 = add_outpt(lp, E[lb].bot);
        add_join(op1, op2, E[lb].top);
      }
      if (E[lb].nael != rb) {
        ix rp = E[rb].pael;
        if (E[rb].outidx >= 0 && E[rp].outidx >= 0 &&
            slopes_equal4(E[rp].cur, E[rp].top, E[rb].cur, E[rb].top)) {
          ix op2 = add_outpt(rp, E[rb].bot);
          add_join(op1, op2, E[rb].top);
        }
        ix e = E[lb].nael;
        if (e != SDC_NIL) {
          while (e != rb) {
            intersect_edges(rb, e, E[lb].cur);
            e = E[e].nael;
          }
        }
      }
    }
  }

  // ------------------------------------------------------------------ horizontals (clipper.cpp:2512-2824)
  SD_HD ix get_maxima_pair(ix e) const {                          // clipper.cpp:2538-2545
    if (pt_eq(E[E[e].next].top, E[e].top) && E[E[e].next].nextlml == SDC_NIL) return E[e].next;
    else if (pt_eq(E[E[e].prev].top, E[e].top) && E[E[e].prev].nextlml == SDC_NIL) return E[e].prev;
    else return SDC_NIL;
  }
  SD_HD ix get_maxima_pair_ex(ix e) const {                       // clipper.cpp:2548-2555
    ix r = get_maxima_pair(e);
    if (r != SDC_NIL && (E[r].nael == E[r].pael && !is_horz(r))) return SDC_NIL;
    return r;
  }
  SD_HD void horz_joins_against_sel(ix horz, ix op1) {
    ix en = sel;
    while (en != SDC_NIL) {
      if (E[en].outidx >= 0 &&
          horz_segments_overlap(E[horz].bot.x, E[horz].top.x, E[en].bot.x, E[en].top.x)) {
        ix op2 = get_last_outpt(en);
        add_join(op2, op1, E[en].top);
      }
      en = E[en].nsel;
    }
  }
  SD_HD_BIG void process_horizontal(ix horz) {
    int dir; i64 hl, hr;
    // GetHorzDirection clipper.cpp:2610-2624
#define SDC_HDIR() do { if (E[horz].bot.x < E[horz].top.x) { hl = E[horz].bot.x; hr = E[horz].top.x; dir = dLeftToRight; } \
                        else { hl = E[horz].top.x; hr = E[horz].bot.x; dir = dRightToLeft; } } while (0)
    SDC_HDIR();
    ix elast = horz, emaxpair = SDC_NIL;
    while (E[elast].nextlml != SDC_NIL && is_horz(E[elast].nextlml)) elast = E[elast].nextlml;
    if (E[elast].nextlml == SDC_NIL) emaxpair = get_maxima_pair(elast);
    ix op1 = SDC_NIL;
    int guard = 0;
    for (;;) {
      bool islast = (horz == elast);
      ix e = (dir == dLeftToRight) ? E[horz].nael : E[horz].pael;
      while (e != SDC_NIL) {
        if (++guard > 100000) { fail(CLIP_OVERFLOW); return; }
        if ((dir == dLeftToRight && E[e].cur.x > hr) || (dir == dRightToLeft && E[e].cur.x < hl)) break;
        if (E[e].cur.x == E[horz].top.x && E[horz].nextlml != SDC_NIL && E[e].dx < E[E[horz].nextlml].dx) break;
        if (E[horz].outidx >= 0) {
          op1 = add_outpt(horz, E[e].cur);
          horz_joins_against_sel(horz, op1);
          add_ghost_join(op1, E[horz].bot);
        }
        if (e == emaxpair && islast) {
          if (E[horz].outidx >= 0) add_local_max_poly(horz, emaxpair, E[horz].top);
          delete_from_ael(horz);
          delete_from_ael(emaxpair);
          return;
        }
        IPt pt; pt.x = E[e].cur.x; pt.y = E[horz].cur.y;
        if (dir == dLeftToRight) intersect_edges(horz, e, pt);
        else intersect_edges(e, horz, pt);
        ix enext = (dir == dLeftToRight) ? E[e].nael : E[e].pael;
        swap_positions_in_ael(horz, e);
        e = enext;
      }
      if (E[horz].nextlml == SDC_NIL || !is_horz(E[horz].nextlml)) break;
      if (!update_edge_into_ael(horz)) return;
      if (E[horz].outidx >= 0) add_outpt(horz, E[horz].bot);
      SDC_HDIR();
    }
#undef SDC_HDIR
    if (E[horz].outidx >= 0 && op1 == SDC_NIL) {
      op1 = get_last_outpt(horz);
      horz_joins_against_sel(horz, op1);
      add_ghost_join(op1, E[horz].top);
    }
    if (E[horz].nextlml != SDC_NIL) {
      if (E[horz].outidx >= 0) {
        op1 = add_outpt(horz, E[horz].top);
        if (!update_edge_into_ael(horz)) return;
        ix ep = E[horz].pael, en = E[horz].nael;
        if (ep != SDC_NIL && E[ep].cur.x == E[horz].bot.x && E[ep].cur.y == E[horz].bot.y &&
            (E[ep].outidx >= 0 && E[ep].cur.y > E[ep].top.y &&
             slopes_equal4_edges(horz, ep))) {
          ix op2 = add_outpt(ep, E[horz].bot);
          add_join(op1, op2, E[horz].top);
        } else if (en != SDC_NIL && E[en].cur.x == E[horz].bot.x && E[en].cur.y == E[horz].bot.y &&
                   E[en].outidx >= 0 && E[en].cur.y > E[en].top.y &&
                   slopes_equal4_edges(horz, en)) {
          ix op2 = add_outpt(en, E[horz].bot);
          add_join(op1, op2, E[horz].top);
        }
      } else update_edge_into_ael(horz);
    } else {
      if (E[horz].outidx >= 0) add_outpt(horz, E[horz].top);
      delete_from_ael(horz);
    }
  }
  // SlopesEqual(const TEdge&, const TEdge&) clipper.cpp:541-551
  SD_HD bool slopes_equal4_edges(ix a, ix b) const {
    return (i64)(E[a].top.y - E[a].bot.y) * (i64)(E[b].top.x - E[b].bot.x) ==
           (i64)(E[a].top.x - E[a].bot.x) * (i64)(E[b].top.y - E[b].bot.y);
  }
  SD_HD void process_horizontals() {
    int guard = 0;
    while (sel != SDC_NIL) {
      if (++guard > 4 * MAXE) { fail(CLIP_OVERFLOW); return; }
      ix h = sel;
      delete_from_sel(h);
      process_horizontal(h);
      if (err) return;
    }
  }

  // ------------------------------------------------------------------ intersections (clipper.cpp:622-690, 2827-2954)
  SD_HD_BIG void intersect_point(ix a, ix b, IPt& ip) {
    const Edge& e1 = E[a]; const Edge& e2 = E[b];
    i64 ipx, ipy;
    double b1, b2;
    if (e1.dx == e2.dx) {
      ipy = e1.cur.y; ipx = top_x(a, ipy);
      ip.x = (int32_t)ipx; ip.y = (int32_t)ipy;
      return;
    } else if (e1.dx == 0) {
      ipx = e1.bot.x;
      if (is_horz(b)) ipy = e2.bot.y;
      else {
        b2 = (double)e2.bot.y - ((double)e2.bot.x / e2.dx);
        ipy = round_haz((double)ipx / e2.dx + b2);
      }
    } else if (e2.dx == 0) {
      ipx = e2.bot.x;
      if (is_horz(a)) ipy = e1.bot.y;
      else {
        b1 = (double)e1.bot.y - ((double)e1.bot.x / e1.dx);
        ipy = round_haz((double)ipx / e1.dx + b1);
      }
    } else {
      b1 = (double)e1.bot.x - (double)e1.bot.y * e1.dx;
      b2 = (double)e2.bot.x - (double)e2.bot.y * e2.dx;
      double q = (b2 - b1) / (e1.dx - e2.dx);
      ipy = round_haz(q);
      if (fabs(e1.dx) < fabs(e2.dx)) ipx = round_haz(e1.dx * q + b1);
      else ipx = round_haz(e2.dx * q + b2);
    }
    if (ipy < e1.top.y || ipy < e2.top.y) {
      if (e1.top.y > e2.top.y) ipy = e1.top.y; else ipy = e2.top.y;
      if (fabs(e1.dx) < fabs(e2.dx)) ipx = top_x(a, ipy); else ipx = top_x(b, ipy);
    }
    if (ipy > e1.cur.y) {
      ipy = e1.cur.y;
      if (fabs(e1.dx) > fabs(e2.dx)) ipx = top_x(b, ipy); else ipx = top_x(a, ipy);
    }
    ip.x = (int32_t)ipx; ip.y = (int32_t)ipy;
  }
  SD_HD_BIG void build_intersect_list(i64 topy) {
    if (ael == SDC_NIL) return;
    ix e = ael;
    sel = e;
    while (e != SDC_NIL) {
      E[e].psel = E[e].pael; E[e].nsel = E[e].nael;
      E[e].cur.x = (int32_t)top_x(e, topy);
      e = E[e].nael;
    }
    bool modified;
    do {
      modified = false;
      e = sel;
      while (E[e].nsel != SDC_NIL) {
        ix en = E[e].nsel;
        if (E[e].cur.x > E[en].cur.x) {
          IPt pt;
          intersect_point(e, en, pt);
          if (pt.y < topy) { pt.x = (int32_t)top_x(e, topy); pt.y = (int32_t)topy; }
          if (nIL >= MAXIL) { fail(CLIP_OVERFLOW); sel = SDC_NIL; return; }
          IL[nIL].e1 = e; IL[nIL].e2 = en; IL[nIL].pt = pt; nIL++; SDC_HW(6, nIL);
          swap_positions_in_sel(e, en);
          modified = true;
        } else e = en;
      }
      if (E[e].psel != SDC_NIL) E[E[e].psel].nsel = SDC_NIL;
      else break;
    } while (modified);
    sel = SDC_NIL;
  }
  SD_HD bool edges_adjacent(const INode& n) const {
    return (E[n.e1].nsel == n.e2) || (E[n.e1].psel == n.e2);
  }
  SD_HD_BIG bool fixup_intersection_order() {
    // CopyAELToSEL clipper.cpp:1929-1939
    ix e = ael; sel = e;
    while (e != SDC_NIL) { E[e].psel = E[e].pael; E[e].nsel = E[e].nael; e = E[e].nael; }
    {
      int32_t keys[MAXIL]; int16_t perm[MAXIL];
      for (int i = 0; i < nIL; ++i) { keys[i] = IL[i].pt.y; perm[i] = (int16_t)i; }
      StdSortDesc<int16_t> s; s.key = keys; s.a = perm; s.sort(nIL);
      // apply permutation in place (cycle-following; MAXIL copies would double the footprint)
      for (int i = 0; i < nIL; ++i) {
        if (perm[i] < 0) continue;
        int j = i; INode tmp = IL[i];
        for (;;) {
          int src = perm[j]; perm[j] = -1 - src;   // mark done (store as negative)
          if (src == i) { IL[j] = tmp; break; }
          IL[j] = IL[src]; j = src;
        }
      }
    }
    for (int i = 0; i < nIL; ++i) {
      if (!edges_adjacent(IL[i])) {
        int j = i + 1;
        while (j < nIL && !edges_adjacent(IL[j])) j++;
        if (j == nIL) return false;
        INode t = IL[i]; IL[i] = IL[j]; IL[j] = t;
      }
      swap_positions_in_sel(IL[i].e1, IL[i].e2);
    }
    return true;
  }
  SD_HD bool process_intersections(i64 topy) {
    if (ael == SDC_NIL) return true;
    nIL = 0;
    build_intersect_list(topy);
    if (err) return false;
    if (nIL == 0) return true;
    if (nIL == 1 || fixup_intersection_order()) {
      for (int i = 0; i < nIL; ++i) {
        intersect_edges(IL[i].e1, IL[i].e2, IL[i].pt);
        swap_positions_in_ael(IL[i].e1, IL[i].e2);
      }
      nIL = 0;
    } else return false;
    sel = SDC_NIL;
    return true;
  }

  // ------------------------------------------------------------------ top of scanbeam (clipper.cpp:2957-3113)
  SD_HD_BIG void do_maxima(ix e) {
    ix emax = get_maxima_pair_ex(e);
    if (emax == SDC_NIL) {
      if (E[e].outidx >= 0) add_outpt(e, E[e].top);
      delete_from_ael(e);
      return;
    }
    ix en = E[e].nael;
    int guard = 0;
    while (en != SDC_NIL && en != emax) {
      if (++guard > 4 * MAXE) { fail(CLIP_OVERFLOW); return; }
      intersect_edges(e, en, E[e].top);
      swap_positions_in_ael(e, en);
      en = E[e].nael;
    }
    if (E[e].outidx == SDC_UNASSIGNED && E[emax].outidx == SDC_UNASSIGNED) {
      delete_from_ael(e); delete_from_ael(emax);
    } else if (E[e].outidx >= 0 && E[emax].outidx >= 0) {
      add_local_max_poly(e, emax, E[e].top);
      delete_from_ael(e); delete_from_ael(emax);
    } else fail(CLIP_FAILED);       // "DoMaxima error" -> caught -> succeeded=false
  }
  SD_HD_BIG void process_edges_at_top_of_scanbeam(i64 topy) {
    ix e = ael;
    int guard = 0;
    while (e != SDC_NIL) {
      if (++guard > 8 * MAXE) { fail(CLIP_OVERFLOW); return; }
      bool ismax = (E[e].top.y == topy && E[e].nextlml == SDC_NIL);
      if (ismax) {
        ix mp = get_maxima_pair_ex(e);
        ismax = (mp == SDC_NIL || !is_horz(mp));
      }
      if (ismax) {
        ix ep = E[e].pael;
        do_maxima(e);
        if (err) return;
        if (ep == SDC_NIL) e = ael; else e = E[ep].nael;
      } else {
        if (E[e].top.y == topy && E[e].nextlml != SDC_NIL && is_horz(E[e].nextlml)) {
          update_edge_into_ael(e);
          if (E[e].outidx >= 0) add_outpt(e, E[e].bot);
          add_edge_to_sel(e);
        } else {
          E[e].cur.x = (int32_t)top_x(e, topy);
          E[e].cur.y = (int32_t)topy;
        }
        e = E[e].nael;
      }
    }
    process_horizontals();
    if (err) return;
    e = ael;
    while (e != SDC_NIL) {
      if (E[e].top.y == topy && E[e].nextlml != SDC_NIL) {
        ix op = SDC_NIL;
        if (E[e].outidx >= 0) op = add_outpt(e, E[e].top);
        if (!update_edge_into_ael(e)) return;
        ix ep = E[e].pael, en = E[e].nael;
        if (ep != SDC_NIL && E[ep].cur.x == E[e].bot.x && E[ep].cur.y == E[e].bot.y && op != SDC_NIL &&
            E[ep].outidx >= 0 && E[ep].cur.y > E[ep].top.y &&
            slopes_equal4(E[e].cur, E[e].top, E[ep].cur, E[ep].top)) {
          ix op2 = add_outpt(ep, E[e].bot);
          add_join(op, op2, E[e].top);
        } else if (en != SDC_NIL && E[en].cur.x == E[e].bot.x && E[en].cur.y == E[e].bot.y && op != SDC_NIL &&
                   E[en].outidx >= 0 && E[en].cur.y > E[en].top.y &&
                   slopes_equal4(E[e].cur, E[e].top, E[en].cur, E[en].top)) {
          ix op2 = add_outpt(en, E[e].bot);
          add_join(op, op2, E[e].top);
        }
      }
      e = E[e].nael;
    }
  }

  // ------------------------------------------------------------------ joins (clipper.cpp:3348-3765)
  SD_HD ix dup_outpt(ix o, bool after) {
    ix r = new_outpt();
    P[r].x = P[o].x; P[r].y = P[o].y; P[r].idx = P[o].idx;
    if (after) { P[r].next = P[o].next; P[r].prev = o; P[P[o].next].prev = r; P[o].next = r; }
    else { P[r].prev = P[o].prev; P[r].next = o; P[P[o].prev].next = r; P[o].prev = r; }
    return r;
  }
  SD_HD bool get_overlap(i64 a1, i64 a2, i64 b1, i64 b2, i64& l, i64& r) const {
#define SDC_MAX(a, b) ((a) < (b) ? (b) : (a))
#define SDC_MIN(a, b) ((b) < (a) ? (b) : (a))
    if (a1 < a2) {
      if (b1 < b2) { l = SDC_MAX(a1, b1); r = SDC_MIN(a2, b2); }
      else { l = SDC_MAX(a1, b2); r = SDC_MIN(a2, b1); }
    } else {
      if (b1 < b2) { l = SDC_MAX(a2, b1); r = SDC_MIN(a1, b2); }
      else { l = SDC_MAX(a2, b2); r = SDC_MIN(a1, b1); }
    }
#undef SDC_MAX
#undef SDC_MIN
    return l < r;
  }
  SD_HD_BIG bool join_horz(ix op1, ix op1b, ix op2, ix op2b, IPt pt, bool discard_left) {
    int dir1 = (P[op1].x > P[op1b].x) ? dRightToLeft : dLeftToRight;
    int dir2 = (P[op2].x > P[op2b].x) ? dRightToLeft : dLeftToRight;
    if (dir1 == dir2) return false;
    int guard = 0;
    if (dir1 == dLeftToRight) {
      while (P[P[op1].next].x <= pt.x && P[P[op1].next].x >= P[op1].x && P[P[op1].next].y == pt.y) {
        op1 = P[op1].next; if (++guard > MAXOP) { fail(CLIP_OVERFLOW); return false; } }
      if (discard_left && (P[op1].x != pt.x)) op1 = P[op1].next;
      op1b = dup_outpt(op1, !discard_left);
      if (P[op1b].x != pt.x || P[op1b].y != pt.y) {
        op1 = op1b; P[op1].x = pt.x; P[op1].y = pt.y; op1b = dup_outpt(op1, !discard_left);
      }
    } else {
      while (P[P[op1].next].x >= pt.x && P[P[op1].next].x <= P[op1].x && P[P[op1].next].y == pt.y) {
        op1 = P[op1].next; if (++guard > MAXOP) { fail(CLIP_OVERFLOW); return false; } }
      if (!discard_left && (P[op1].x != pt.x)) op1 = P[op1].next;
      op1b = dup_outpt(op1, discard_left);
      if (P[op1b].x != pt.x || P[op1b].y != pt.y) {
        op1 = op1b; P[op1].x = pt.x; P[op1].y = pt.y; op1b = dup_outpt(op1, discard_left);
      }
    }
    guard = 0;
    if (dir2 == dLeftToRight) {
      while (P[P[op2].next].x <= pt.x && P[P[op2].next].x >= P[op2].x && P[P[op2].next].y == pt.y) {
        op2 = P[op2].next; if (++guard > MAXOP) { fail(CLIP_OVERFLOW); return false; } }
      if (discard_left && (P[op2].x != pt.x)) op2 = P[op2].next;
      op2b = dup_outpt(op2, !discard_left);
      if (P[op2b].x != pt.x || P[op2b].y != pt.y) {
        op2 = op2b; P[op2].x = pt.x; P[op2].y = pt.y; op2b = dup_outpt(op2, !discard_left);
      }
    } else {
      while (P[P[op2].next].x >= pt.x && P[P[op2].next].x <= P[op2].x && P[P[op2].next].y == pt.y) {
        op2 = P[op2].next; if (++guard > MAXOP) { fail(CLIP_OVERFLOW); return false; } }
      if (!discard_left && (P[op2].x != pt.x)) op2 = P[op2].next;
      op2b = dup_outpt(op2, discard_left);
      if (P[op2b].x != pt.x || P[op2b].y != pt.y) {
        op2 = op2b; P[op2].x = pt.x; P[op2].y = pt.y; op2b = dup_outpt(op2, discard_left);
      }
    }
    if ((dir1 == dLeftToRight) == discard_left) {
      P[op1].prev = op2; P[op2].next = op1; P[op1b].next = op2b; P[op2b].prev = op1b;
    } else {
      P[op1].next = op2; P[op2].prev = op1; P[op1b].prev = op2b; P[op2b].next = op1b;
    }
    return true;
  }
  SD_HD_BIG bool join_points(Join& j, ix r1, ix r2) {
    ix op1 = j.op1, op1b, op2 = j.op2, op2b;
    bool horizontal = (P[j.op1].y == j.off.y);
    if (horizontal && pt_eq(j.off, opt(j.op1)) && pt_eq(j.off, opt(j.op2))) {
      // strictly-simple style join (all three points coincide)
      if (r1 != r2) return false;
      op1b = P[j.op1].next;
      while (op1b != op1 && pt_eq(opt(op1b), j.off)) op1b = P[op1b].next;
      bool rev1 = (P[op1b].y > j.off.y);
      op2b = P[j.op2].next;
      while (op2b != op2 && pt_eq(opt(op2b), j.off)) op2b = P[op2b].next;
      bool rev2 = (P[op2b].y > j.off.y);
      if (rev1 == rev2) return false;
      if (rev1) {
        op1b = dup_outpt(op1, false); op2b = dup_outpt(op2, true);
        P[op1].prev = op2; P[op2].next = op1; P[op1b].next = op2b; P[op2b].prev = op1b;
        j.op1 = op1; j.op2 = op1b; return true;
      } else {
        op1b = dup_outpt(op1, true); op2b = dup_outpt(op2, false);
        P[op1].next = op2; P[op2].prev = op1; P[op1b].prev = op2b; P[op2b].next = op1b;
        j.op1 = op1; j.op2 = op1b; return true;
      }
    } else if (horizontal) {
      op1b = op1;
      while (P[P[op1].prev].y == P[op1].y && P[op1].prev != op1b && P[op1].prev != op2) op1 = P[op1].prev;
      while (P[P[op1b].next].y == P[op1b].y && P[op1b].next != op1 && P[op1b].next != op2) op1b = P[op1b].next;
      if (P[op1b].next == op1 || P[op1b].next == op2) return false;
      op2b = op2;
      while (P[P[op2].prev].y == P[op2].y && P[op2].prev != op2b && P[op2].prev != op1b) op2 = P[op2].prev;
      while (P[P[op2b].next].y == P[op2b].y && P[op2b].next != op2 && P[op2b].next != op1) op2b = P[op2b].next;
      if (P[op2b].next == op2 || P[op2b].next == op1) return false;
      i64 l, r;
      if (!get_overlap(P[op1].x, P[op1b].x, P[op2].x, P[op2b].x, l, r)) return false;
      IPt pt; bool discard_left;
      if (P[op1].x >= l && P[op1].x <= r) { pt = opt(op1); discard_left = (P[op1].x > P[op1b].x); }
      else if (P[op2].x >= l && P[op2].x <= r) { pt = opt(op2); discard_left = (P[op2].x > P[op2b].x); }
      else if (P[op1b].x >= l && P[op1b].x <= r) { pt = opt(op1b); discard_left = P[op1b].x > P[op1].x; }
      else { pt = opt(op2b); discard_left = (P[op2b].x > P[op2].x); }
      j.op1 = op1; j.op2 = op2;
      return join_horz(op1, op1b, op2, op2b, pt, discard_left);
    } else {
      op1b = P[op1].next;
      while (pt_eq(opt(op1b), opt(op1)) && op1b != op1) op1b = P[op1b].next;
      bool rev1 = ((P[op1b].y > P[op1].y) || !slopes_equal3(opt(op1), opt(op1b), j.off));
      if (rev1) {
        op1b = P[op1].prev;
        while (pt_eq(opt(op1b), opt(op1)) && op1b != op1) op1b = P[op1b].prev;
        if ((P[op1b].y > P[op1].y) || !slopes_equal3(opt(op1), opt(op1b), j.off)) return false;
      }
      op2b = P[op2].next;
      while (pt_eq(opt(op2b), opt(op2)) && op2b != op2) op2b = P[op2b].next;
      bool rev2 = ((P[op2b].y > P[op2].y) || !slopes_equal3(opt(op2), opt(op2b), j.off));
      if (rev2) {
        op2b = P[op2].prev;
        while (pt_eq(opt(op2b), opt(op2)) && op2b != op2) op2b = P[op2b].prev;
        if ((P[op2b].y > P[op2].y) || !slopes_equal3(opt(op2), opt(op2b), j.off)) return false;
      }
      if (op1b == op1 || op2b == op2 || op1b == op2b || ((r1 == r2) && (rev1 == rev2))) return false;
      if (rev1) {
        op1b = dup_outpt(op1, false); op2b = dup_outpt(op2, true);
        P[op1].prev = op2; P[op2].next = op1; P[op1b].next = op2b; P[op2b].prev = op1b;
        j.op1 = op1; j.op2 = op1b; return true;
      } else {
        op1b = dup_outpt(op1, true); op2b = dup_outpt(op2, false);
        P[op1].next = op2; P[op2].prev = op1; P[op1b].prev = op2b; P[op2b].next = op1b;
        j.op1 = op1; j.op2 = op1b; return true;
      }
    }
  }
  // clipper.cpp:484-523; 0 outside, +1 inside, -1 on boundary
  SD_HD_BIG int point_in_polygon(IPt pt, ix op) const {
    int result = 0;
    ix start = op;
    for (;;) {
      ix nx = P[op].next;
      if (P[nx].y == pt.y) {
        if ((P[nx].x == pt.x) || (P[op].y == pt.y && ((P[nx].x > pt.x) == (P[op].x < pt.x)))) return -1;
      }
      if ((P[op].y < pt.y) != (P[nx].y < pt.y)) {
        if (P[op].x >= pt.x) {
          if (P[nx].x > pt.x) result = 1 - result;
          else {
            double d = (double)((i64)P[op].x - pt.x) * (double)((i64)P[nx].y - pt.y) -
                       (double)((i64)P[nx].x - pt.x) * (double)((i64)P[op].y - pt.y);
            if (!d) return -1;
            if ((d > 0) == (P[nx].y > P[op].y)) result = 1 - result;
          }
        } else {
          if (P[nx].x > pt.x) {
            double d = (double)((i64)P[op].x - pt.x) * (double)((i64)P[nx].y - pt.y) -
                       (double)((i64)P[nx].x - pt.x) * (double)((i64)P[op].y - pt.y);
            if (!d) return -1;
            if ((d > 0) == (P[nx].y > P[op].y)) result = 1 - result;
          }
        }
      }
      op = nx;
      if (start == op) break;
    }
    return result;
  }
  SD_HD bool poly2_contains_poly1(ix o1, ix o2) const {           // clipper.cpp:526-538
    ix op = o1;
    do {
      int res = point_in_polygon(opt(op), o2);
      if (res >= 0) return res > 0;
      op = P[op].next;
    } while (op != o1);
    return true;
  }
  SD_HD_BIG void join_common_edges() {
    for (int i = 0; i < nJ; ++i) {
      Join& j = J[i];
      ix r1 = get_outrec(P[j.op1].idx);
      ix r2 = get_outrec(P[j.op2].idx);
      if (R[r1].pts == SDC_NIL || R[r2].pts == SDC_NIL) continue;
      ix holerec;
      if (r1 == r2) holerec = r1;
      else if (outrec1_right_of_outrec2(r1, r2)) holerec = r2;
      else if (outrec1_right_of_outrec2(r2, r1)) holerec = r1;
      else holerec = get_lowermost_rec(r1, r2);
      if (!join_points(j, r1, r2)) continue;
      if (err) return;
      if (r1 == r2) {
        R[r1].pts = j.op1; R[r1].bottom = SDC_NIL;
        r2 = create_outrec();
        R[r2].pts = j.op2;
        { ix op = R[r2].pts; int guard = 0;
          do { P[op].idx = R[r2].idx; op = P[op].prev; if (++guard > MAXOP) { fail(CLIP_OVERFLOW); return; } } while (op != R[r2].pts); }
        if (poly2_contains_poly1(R[r2].pts, R[r1].pts)) {
          R[r2].ishole = (int8_t)!R[r1].ishole; R[r2].firstleft = r1;
          if ((R[r2].ishole != 0) == (area_op(R[r2].pts) > 0)) reverse_poly_pt_links(R[r2].pts);
        } else if (poly2_contains_poly1(R[r1].pts, R[r2].pts)) {
          R[r2].ishole = R[r1].ishole; R[r1].ishole = (int8_t)!R[r2].ishole;
          R[r2].firstleft = R[r1].firstleft; R[r1].firstleft = r2;
          if ((R[r1].ishole != 0) == (area_op(R[r1].pts) > 0)) reverse_poly_pt_links(R[r1].pts);
        } else {
          R[r2].ishole = R[r1].ishole; R[r2].firstleft = R[r1].firstleft;
        }
      } else {
        R[r2].pts = SDC_NIL; R[r2].bottom = SDC_NIL; R[r2].idx = R[r1].idx;
        R[r1].ishole = R[holerec].ishole;
        if (holerec == r2) R[r1].firstleft = R[r2].firstleft;
        R[r2].firstleft = r1;
      }
    }
  }
  SD_HD_BIG void fixup_out_polygon(ix orec) {                         // clipper.cpp:3143-3181
    ix lastok = SDC_NIL;
    R[orec].bottom = SDC_NIL;
    ix pp = R[orec].pts;
    int guard = 0;
    for (;;) {
      if (++guard > 4 * MAXOP) { fail(CLIP_OVERFLOW); return; }
      if (P[pp].prev == pp || P[pp].prev == P[pp].next) { R[orec].pts = SDC_NIL; return; }
      if (pt_eq(opt(pp), opt(P[pp].next)) || pt_eq(opt(pp), opt(P[pp].prev)) ||
          slopes_equal3(opt(P[pp].prev), opt(pp), opt(P[pp].next))) {
        lastok = SDC_NIL;
        P[P[pp].prev].next = P[pp].next;
        P[P[pp].next].prev = P[pp].prev;
        pp = P[pp].prev;
      } else if (pp == lastok) break;
      else { if (lastok == SDC_NIL) lastok = pp; pp = P[pp].next; }
    }
    R[orec].pts = pp;
  }

  // ------------------------------------------------------------------ Execute (clipper.cpp:1508-1621)
  SD_HD bool execute() {
    reset();
    sel = SDC_NIL;
    nR = 0; nP = 0; nJ = 0; nGJ = 0; nIL = 0;
    bool ok = true;
    i64 boty, topy = 0;
    if (!sb_pop(boty)) return false;
    insert_local_minima_into_ael(boty);
    int guard = 0;
    while (!err) {
      bool popped = sb_pop(topy);
      if (!popped && !(curLM < nLM)) break;
      if (++guard > 8 * MAXE) { fail(CLIP_OVERFLOW); break; }
      process_horizontals();
      if (err) break;
      nGJ = 0;
      if (!process_intersections(topy)) { ok = false; break; }
      process_edges_at_top_of_scanbeam(topy);
      if (err) break;
      boty = topy;
      insert_local_minima_into_ael(boty);
    }
    if (err) ok = false;
    if (ok) {
      for (int i = 0; i < nR; ++i) {
        if (R[i].pts == SDC_NIL) continue;
        if ((R[i].ishole != 0) == (area_op(R[i].pts) > 0)) reverse_poly_pt_links(R[i].pts);
      }
      if (nJ > 0) join_common_edges();
      for (int i = 0; i < nR; ++i) {
        if (R[i].pts == SDC_NIL) continue;
        fixup_out_polygon((ix)i);
      }
      if (err) ok = false;
    }
    return ok;
  }

  // number of result paths / enumerate in BuildResult order (clipper.cpp:3199-3217)
  SD_HD int point_count(ix p) const {
    if (p == SDC_NIL) return 0;
    int c = 0; ix q = p;
    do { c++; q = P[q].next; } while (q != p && c <= MAXOP);
    return c;
  }

  // area_from_path summed over result paths (stardist2d.cpp:128-138,161-164):
  // float accumulator over int64 cross products, vertices in BuildResult order.
  SD_HD float result_area() const {
    float total = 0;
    for (int i = 0; i < nR; ++i) {
      if (R[i].pts == SDC_NIL) continue;
      ix p0 = P[R[i].pts].prev;
      int cnt = point_count(p0);
      if (cnt < 2) continue;
      float area = 0;
      ix p = p0;
      for (int k = 0; k < cnt; ++k) {
        ix q = (k == cnt - 1) ? p0 : P[p].prev;
        i64 cr = (i64)P[p].x * (i64)P[q].y - (i64)P[p].y * (i64)P[q].x;
        area = area + (float)cr;
        p = P[p].prev;
      }
      // area = 0.5 * abs(area): double product, stored back to float
      area = (float)(0.5 * (double)fabsf(area));
      total = total + area;
    }
    return total;
  }
};

// Full pair test. ax/ay = polygon i (added first, as ptClip), bx/by = polygon j (ptSubject).
// Returns the intersection area exactly as poly_intersection_area() would (0 when Execute fails
// or a path is rejected); *status receives a ClipErr.
// vertex sources
struct SplitXY {            // separate x[] / y[] arrays
  const int32_t* xs; const int32_t* ys;
  SD_HD int32_t x(int i) const { return xs[i]; }
  SD_HD int32_t y(int i) const { return ys[i]; }
};
struct InterleavedXY {      // (x,y) pairs
  const int32_t* xy;
  SD_HD int32_t x(int i) const { return xy[2 * i]; }
  SD_HD int32_t y(int i) const { return xy[2 * i + 1]; }
};

// Full pair test. a = polygon i (added first, as ptClip), b = polygon j (ptSubject).
// Returns the intersection area exactly as poly_intersection_area() would (0 when Execute fails
// or a path is rejected); *status receives a ClipErr.
template <int NV, int SC, typename VA, typename VB>
SD_HD float clip_intersection_area(const VA& a, const VB& b, int n, ClipSweep<NV, SC>& S, int* status) {
  S.init();
  S.add_path(a, n, ptClip);
  S.add_path(b, n, ptSubject);
  float area = 0.f;
  if (!S.err) {
    bool ok = S.execute();
    if (ok) area = S.result_area();
  }
  *status = S.err;
  if (S.err == CLIP_OVERFLOW) return 0.f;
  return area;
}

}  // namespace sdclip
