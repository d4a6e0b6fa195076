// Host-side index sampling of the ASLSD estimator, bit-compatible with numpy's Generator(PCG64).
//
// The stochastic estimates match the reference only if its numpy Generator is consumed draw for
// draw (BASELINE.json north_star: "host-generated stratum indices so stochastic estimates match
// the reference exactly").  The reference samples with one rng.choice call per lag
// (utilities/useful_statistics.py:38-109): tens of thousands of tiny numpy calls per iteration.
// This file restates the three pieces of numpy those calls go through - the PCG64 XSL-RR 128/64
// generator with its 32-bit output buffer, Lemire's bounded integers on 32-bit draws, and
// Generator.choice(replace=False) (Floyd's algorithm with an open-addressing set, tail shuffle
// for large samples, final shuffle) - so that a whole term is sampled in one native call that
// starts from the Generator's state and hands the advanced state back.  No device code here.
// Bit-compatibility with the installed numpy is pinned by tests/test_sampling_stream.py.
#include <stdint.h>

#include <vector>

#include "../../include/aslsd_b200.h"

typedef unsigned __int128 u128;

namespace {

struct Pcg {
  u128 state, inc;
  int has_uint32;
  uint32_t uinteger;
  explicit Pcg(const aslsd_pcg64& g)
      : state(((u128)g.state_hi << 64) | g.state_lo), inc(((u128)g.inc_hi << 64) | g.inc_lo),
        has_uint32(g.has_uint32), uinteger(g.uinteger) {}
  void store(aslsd_pcg64* g) const {
    g->state_hi = (uint64_t)(state >> 64);
    g->state_lo = (uint64_t)state;
    g->has_uint32 = has_uint32;
    g->uinteger = uinteger;
  }
  uint64_t next64() {
    const u128 mult = ((u128)0x2360ED051FC65DA4ULL << 64) | 0x4385DF649FCCF645ULL;
    state = state * mult + inc;
    const uint64_t hi = (uint64_t)(state >> 64), lo = (uint64_t)state;
    const uint64_t x = hi ^ lo;
    const unsigned rot = (unsigned)(hi >> 58);
    return (x >> rot) | (x << ((64 - rot) & 63));
  }
  uint32_t next32() {
    if (has_uint32) {
      has_uint32 = 0;
      return uinteger;
    }
    const uint64_t v = next64();
    has_uint32 = 1;
    uinteger = (uint32_t)(v >> 32);
    return (uint32_t)(v & 0xffffffffULL);
  }
  // random_bounded_uint64(bitgen, 0, rng, 0, use_masked = false): uniform on [0, rng]
  uint64_t bounded(uint64_t rng) {
    if (rng == 0) return 0;
    if (rng <= 0xFFFFFFFFULL) {
      if (rng == 0xFFFFFFFFULL) return next32();
      const uint32_t rng_excl = (uint32_t)rng + 1u;
      uint64_t m = (uint64_t)next32() * (uint64_t)rng_excl;
      uint32_t leftover = (uint32_t)(m & 0xFFFFFFFFULL);
      if (leftover < rng_excl) {
        const uint32_t threshold = (uint32_t)((0xFFFFFFFFULL - rng) % rng_excl);
        while (leftover < threshold) {
          m = (uint64_t)next32() * (uint64_t)rng_excl;
          leftover = (uint32_t)(m & 0xFFFFFFFFULL);
        }
      }
      return m >> 32;
    }
    if (rng == 0xFFFFFFFFFFFFFFFFULL) return next64();
    const uint64_t rng_excl = rng + 1;
    u128 m = (u128)next64() * (u128)rng_excl;
    uint64_t leftover = (uint64_t)m;
    if (leftover < rng_excl) {
      const uint64_t threshold = (0xFFFFFFFFFFFFFFFFULL - rng) % rng_excl;
      while (leftover < threshold) {
        m = (u128)next64() * (u128)rng_excl;
        leftover = (uint64_t)m;
      }
    }
    return (uint64_t)(m >> 64);
  }
};

// _shuffle_int(bitgen, n, first, data)
void shuffle_int(Pcg& g, int64_t n, int64_t first, int64_t* data) {
  for (int64_t i = n - 1; i >= first; --i) {
    const int64_t j = (int64_t)g.bounded((uint64_t)i);
    const int64_t tmp = data[j];
    data[j] = data[i];
    data[i] = tmp;
  }
}

// Generator.choice(pop, size, replace=False, shuffle=True) without p
int choice_no_replace(Pcg& g, int64_t pop, int64_t size, int64_t* out, std::vector<int64_t>& scratch,
                      std::vector<uint64_t>& set) {
  if (size < 0 || pop < 0 || size > pop) return ASLSD_ERR_ARG;
  if (size == 0) return ASLSD_OK;
  if (size == 1) {
    // Floyd's algorithm with one element: a single bounded draw on [0, pop - 1], nothing to shuffle
    out[0] = (int64_t)g.bounded((uint64_t)(pop - 1));
    return ASLSD_OK;
  }
  if (pop > 10000 && size > pop / 50) {
    // tail shuffle of arange(pop): i from pop - 1 down to first, j uniform on [0, i], swap
    int64_t first = pop - size;
    if (first < 1) first = 1;
    if (size * 4 >= pop) {
      scratch.resize((size_t)pop);
      for (int64_t q = 0; q < pop; ++q) scratch[(size_t)q] = q;
      shuffle_int(g, pop, first, scratch.data());
      for (int64_t q = 0; q < size; ++q) out[q] = scratch[(size_t)(pop - size + q)];
      return ASLSD_OK;
    }
    // The same swaps without materialising arange(pop): position i is final once its own step
    // is done (later steps only touch positions below it), so only the displaced values at the
    // swap targets j are kept, in an open-addressing map (absent key = identity).
    uint64_t mask = (uint64_t)(2 * size);
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16; mask |= mask >> 32;
    set.assign(2 * ((size_t)mask + 1), ~(uint64_t)0);   // (key, value) pairs
    auto slot = [&](uint64_t key) {
      uint64_t loc = (key * 0x9E3779B97F4A7C15ULL >> 20) & mask;
      while (set[2 * loc] != ~(uint64_t)0 && set[2 * loc] != key) loc = (loc + 1) & mask;
      return loc;
    };
    for (int64_t i = pop - 1; i >= first; --i) {
      const uint64_t li = slot((uint64_t)i);
      const int64_t vi = set[2 * li] == (uint64_t)i ? (int64_t)set[2 * li + 1] : i;
      const int64_t j = (int64_t)g.bounded((uint64_t)i);
      int64_t vj = vi;
      if (j != i) {
        const uint64_t lj = slot((uint64_t)j);
        vj = set[2 * lj] == (uint64_t)j ? (int64_t)set[2 * lj + 1] : j;
        set[2 * lj] = (uint64_t)j;
        set[2 * lj + 1] = (uint64_t)vi;
      }
      out[i - (pop - size)] = vj;
    }
    return ASLSD_OK;
  }
  // Floyd's algorithm; set size = smallest power of two above 1.2 * size
  uint64_t mask = (uint64_t)(1.2 * (double)size);
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16; mask |= mask >> 32;
  set.assign((size_t)mask + 1, ~(uint64_t)0);
  for (int64_t j = pop - size; j < pop; ++j) {
    const uint64_t val = g.bounded((uint64_t)j);
    uint64_t loc = val & mask;
    while (set[loc] != ~(uint64_t)0 && set[loc] != val) loc = (loc + 1) & mask;
    if (set[loc] == ~(uint64_t)0) {
      set[loc] = val;
      out[j - pop + size] = (int64_t)val;
    } else {
      loc = (uint64_t)j & mask;
      while (set[loc] != ~(uint64_t)0) loc = (loc + 1) & mask;
      set[loc] = (uint64_t)j;
      out[j - pop + size] = j;
    }
  }
  shuffle_int(g, size, 1, out);
  return ASLSD_OK;
}

}  // namespace

extern "C" int aslsd_rng_choice(aslsd_pcg64* gen, int64_t pop, int64_t size, int replace, int64_t* out) {
  if (!gen || (!out && size > 0)) return ASLSD_ERR_ARG;
  Pcg g(*gen);
  int rc = ASLSD_OK;
  if (replace) {
    if (pop <= 0 && size > 0) return ASLSD_ERR_ARG;
    for (int64_t q = 0; q < size; ++q) out[q] = (int64_t)g.bounded((uint64_t)(pop - 1));
  } else {
    std::vector<int64_t> scratch;
    std::vector<uint64_t> set;
    rc = choice_no_replace(g, pop, size, out, scratch, set);
  }
  if (rc == ASLSD_OK) g.store(gen);
  return rc;
}

// All strata of one single-sum term: sample_single_ix_stratified (useful_statistics.py:38-60),
// lo[s] + choice(hi[s] - lo[s] + 1, n_samples[s], replace=False) for every stratum in turn.
extern "C" int aslsd_sample_single_ix(aslsd_pcg64* gen, int64_t n_strata, const int64_t* lo, const int64_t* hi,
                                      const int64_t* n_samples, int64_t* out) {
  if (!gen || n_strata < 0 || (n_strata > 0 && (!lo || !hi || !n_samples))) return ASLSD_ERR_ARG;
  Pcg g(*gen);
  std::vector<int64_t> scratch;
  std::vector<uint64_t> set;
  int64_t pos = 0;
  for (int64_t s = 0; s < n_strata; ++s) {
    const int64_t ns = n_samples[s], pop = hi[s] - lo[s] + 1;
    if (ns < 0 || (ns > 0 && !out)) return ASLSD_ERR_ARG;
    const int rc = choice_no_replace(g, pop, ns, out + pos, scratch, set);
    if (rc) return rc;
    for (int64_t q = 0; q < ns; ++q) out[pos + q] += lo[s];
    pos += ns;
  }
  g.store(gen);
  return ASLSD_OK;
}

// All strata of one double-sum term: sample_double_ix_stratified (useful_statistics.py:112-127)
// over sample_double_ix_stratum (:73-109).  varpi_ij = varpi[i][j], kappa_ji = kappa[j][i].
// Stratum s writes n_samples[s] pairs at m_out / n_out + sum(n_samples[:s]).
extern "C" int aslsd_sample_double_ix(aslsd_pcg64* gen, int64_t n_strata, const int64_t* h_inf, const int64_t* h_sup,
                                      const int64_t* n_samples, const int64_t* varpi_ij, const int64_t* kappa_ji,
                                      int64_t n_events_i, int64_t* m_out, int64_t* n_out) {
  if (!gen || n_strata < 0 || (n_strata > 0 && (!h_inf || !h_sup || !n_samples || !varpi_ij || !kappa_ji)))
    return ASLSD_ERR_ARG;
  Pcg g(*gen);
  std::vector<int64_t> scratch, lags, order, count, ckey, slot_of;
  std::vector<uint64_t> set;
  int64_t pos = 0;
  for (int64_t s = 0; s < n_strata; ++s) {
    const int64_t ns = n_samples[s];
    if (ns < 0) return ASLSD_ERR_ARG;
    int64_t* m = m_out + pos;
    int64_t* n = n_out + pos;
    if (h_inf[s] == h_sup[s]) {
      const int64_t h = h_inf[s], first = varpi_ij[h];
      int rc = choice_no_replace(g, n_events_i - first, ns, m, scratch, set);
      if (rc) return rc;
      for (int64_t q = 0; q < ns; ++q) {
        m[q] += first;
        __builtin_prefetch(&kappa_ji[m[q]]);
      }
      for (int64_t q = 0; q < ns; ++q) n[q] = kappa_ji[m[q]] - (h - 1);
    } else {
      // lags with replacement, then one choice per distinct lag in order of first occurrence
      // (collections.Counter keeps insertion order)
      const int64_t width = h_sup[s] + 1 - h_inf[s];
      if (width <= 0 && ns > 0) return ASLSD_ERR_ARG;
      lags.resize((size_t)ns);
      for (int64_t q = 0; q < ns; ++q) lags[(size_t)q] = (int64_t)g.bounded((uint64_t)(width - 1));
      // multiplicity of every distinct lag, in order of first occurrence: a small open-addressing
      // table (a dense array over the lag range would be megabytes for the outer strata)
      uint64_t cmask = (uint64_t)(2 * ns + 1);
      cmask |= cmask >> 1; cmask |= cmask >> 2; cmask |= cmask >> 4; cmask |= cmask >> 8; cmask |= cmask >> 16;
      cmask |= cmask >> 32;
      ckey.assign((size_t)cmask + 1, -1);
      count.assign((size_t)cmask + 1, 0);
      order.clear();
      slot_of.clear();
      for (int64_t q = 0; q < ns; ++q) {
        const int64_t lag = lags[(size_t)q];
        uint64_t loc = ((uint64_t)lag * 0x9E3779B97F4A7C15ULL >> 24) & cmask;
        while (ckey[loc] != -1 && ckey[loc] != lag) loc = (loc + 1) & cmask;
        if (ckey[loc] == -1) {
          ckey[loc] = lag;
          order.push_back(lag);
          slot_of.push_back((int64_t)loc);
        }
        ++count[loc];
      }
      // varpi[h] and kappa[m] are scattered reads into tables of the size of the event streams:
      // prefetched ahead of their use
      for (size_t o = 0; o < order.size(); ++o) __builtin_prefetch(&varpi_ij[h_inf[s] + order[o]]);
      int64_t w = 0;
      for (size_t o = 0; o < order.size(); ++o) {
        const int64_t rel = order[o];
        const int64_t h = h_inf[s] + rel, n_h = count[(size_t)slot_of[o]];
        const int64_t first = varpi_ij[h], pop = n_events_i - first;
        if (n_h <= pop) {
          int rc = choice_no_replace(g, pop, n_h, m + w, scratch, set);
          if (rc) return rc;
        } else {
          if (pop <= 0) return ASLSD_ERR_ARG;
          for (int64_t q = 0; q < n_h; ++q) m[w + q] = (int64_t)g.bounded((uint64_t)(pop - 1));
        }
        for (int64_t q = 0; q < n_h; ++q) {
          m[w + q] += first;
          __builtin_prefetch(&kappa_ji[m[w + q]]);
          n[w + q] = h;          // the lag, until kappa has arrived
        }
        w += n_h;
      }
      for (int64_t q = 0; q < ns; ++q) n[q] = kappa_ji[m[q]] - n[q] + 1;
    }
    pos += ns;
  }
  g.store(gen);
  return ASLSD_OK;
}
