// The per-entry functions of vstree_b200/csrc/childtab_core.cuh (the code the child-table kernels run per thread)
// compiled for the HOST, so that tests/test_childtab_model.py can compare them with the oracle's stack walks without a
// GPU.  Test infrastructure: the product never builds or loads this file.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#define CLD_NEAR32 1  // the experiment variant of the bounded scan is modelled too (near32 != 0 below)
#include "../../vstree_b200/csrc/childtab_core.cuh"

namespace
{
struct HostPyramid
{
  std::vector<std::vector<uint32_t>> lev;
  cld::Pyramid py;
  HostPyramid(const uint32_t *L, uint64_t N)
  {
    lev.emplace_back(L, L + N);
    while (lev.back().size() > cld::FAN && (int) lev.size() < cld::MAX_LEVELS)
    {
      const std::vector<uint32_t> &in = lev.back();
      std::vector<uint32_t> out((in.size() + cld::FAN - 1) / cld::FAN, 0xffffffffu);
      for (size_t i = 0; i < in.size(); i++)
      {
        if (in[i] < out[i / cld::FAN])
        {
          out[i / cld::FAN] = in[i];
        }
      }
      lev.push_back(out);
    }
    memset(&py, 0, sizeof py);
    py.top = (int) lev.size() - 1;
    for (size_t k = 0; k < lev.size(); k++)
    {
      py.level[k] = lev[k].data();
      py.size[k] = lev[k].size();
    }
  }
};
}  // namespace

extern "C" {

// L has N = n + 1 entries (L[0] = 0); cld gets 3 N bytes, cld1 N bytes
void cld_model_tables(const uint32_t *L, uint64_t N, uint32_t threads, uint32_t near, int near32, uint8_t *cldtab,
                      uint8_t *cld1, uint64_t *near_count)
{
  HostPyramid h(L, N);
  // as child_kernel runs it: blocks of `threads` entries with a window of `near` entries on either side; the bounded
  // scan first, the pyramid search for what it leaves open (near == 0: the pyramid search for every entry)
  std::vector<uint32_t> win(threads + 2 * near);
  for (uint64_t i = 0; i < N; i++)
  {
    cld::Child c;
    bool settled = false;
    if (near > 0)
    {
      const uint64_t first = i / threads * threads;
      const uint64_t w0 = first >= near ? first - near : 0;
      const uint64_t wend = first + threads + near < N ? first + threads + near : N;
      for (uint64_t t = w0; t < wend; t++)
      {
        win[t - w0] = L[t];
      }
      settled = near32 ? cld::child_entry_near32(win.data(), (uint32_t) (wend - w0), wend == N, (uint32_t) (i - w0), near, c)
                       : cld::child_entry_near(win.data(), w0, wend - w0, N, i, near, c);
      if (settled && near_count != nullptr)
      {
        ++*near_count;
      }
    }
    if (!settled)
    {
      c = cld::child_entry(h.py, i);
    }
    cldtab[3 * i] = c.up;
    cldtab[3 * i + 1] = c.down;
    cldtab[3 * i + 2] = c.nextl;
  }
  for (uint64_t i = 0; i < N; i++)
  {
    cld1[i] = cld::cld1_entry(cldtab, L, i, N - 1);
  }
}

// the three searches on their own: q[t] = (kind, a, b); kind 0 next_le(a, value b), 1 prev_le(a, value b),
// 2 leftmost_min(a, b)
void cld_model_queries(const uint32_t *L, uint64_t N, const uint64_t *q, uint64_t nq, uint64_t *answer)
{
  HostPyramid h(L, N);
  for (uint64_t t = 0; t < nq; t++)
  {
    const uint64_t kind = q[3 * t], a = q[3 * t + 1], b = q[3 * t + 2];
    answer[t] = kind == 0 ? cld::next_le(h.py, a, (uint32_t) b)
                          : kind == 1 ? cld::prev_le(h.py, a, (uint32_t) b) : cld::leftmost_min(h.py, a, b);
  }
}
}
