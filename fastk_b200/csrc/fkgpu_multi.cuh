/*  fkgpu_multi.cuh -- the multi-GPU count behind the C ABI: one context per GPU, one NCCL communicator,
 *  the whole exchange inside the library.  Included by fkgpu_api.cu.
 *
 *  Reads are data-parallel across the ranks.  Every rank scans ITS reads into 8-byte super-mer records that point into the
 *  concatenation of all ranks' read streams, partitioned by the top bucket bits.  Minimizer buckets are sharded in contiguous
 *  ranges (cumulative-threshold rule of the reference's thread split, MSDsort.c:330-352, on the all-reduced histogram); ONE
 *  grouped ncclSend/ncclRecv all-to-all moves the records, a second their 32-byte base strings.  Every instance of a canonical
 *  k-mer lives in one bucket, so the owner counts its buckets completely on chip -- no count ever merges across ranks.  Only
 *  when a table is wanted the distinct entries take a second all-to-all, by key prefix, and are put in key order locally:
 *  rank order == key order, the global table is the rank-ordered concatenation (what Merge_Tables would otherwise merge,
 *  table.c:346-533).  Host round trips per count: the sizes of the two exchanges (one all-gather each), nothing per stage.
 *
 *  NCCL is bound at run time (dlopen of libnccl.so.2): a single-GPU user of the library needs no NCCL.                   */
#pragma once
#include <dlfcn.h>

namespace fkmg {

/* the handful of NCCL declarations used (ABI of nccl.h 2.x) */
typedef struct { char internal[128]; } UniqueId;
typedef void *Comm;
enum { kUint8 = 1, kInt64 = 4, kUint64 = 5, kSum = 0 };

struct Api
  { void *lib = nullptr;
    int (*GetUniqueId)(UniqueId *) = nullptr;
    int (*CommInitRank)(Comm *, int, UniqueId, int) = nullptr;
    int (*CommDestroy)(Comm) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void *, size_t, int, int, Comm, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, Comm, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, Comm, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, Comm, cudaStream_t) = nullptr;
  };

static Api *api()
{ static Api a;
  static std::once_flag once;
  std::call_once(once,[]()
    { const char *names[] = { getenv("FKGPU_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
      for (const char *n : names)
        { if (n == NULL) continue;
          a.lib = dlopen(n,RTLD_NOW | RTLD_GLOBAL);
          if (a.lib) break;
        }
      if (a.lib == NULL) return;
#define FKMG_SYM(f) *(void **) &a.f = dlsym(a.lib,"nccl" #f)
      FKMG_SYM(GetUniqueId); FKMG_SYM(CommInitRank); FKMG_SYM(CommDestroy); FKMG_SYM(GetErrorString); FKMG_SYM(GroupStart);
      FKMG_SYM(GroupEnd); FKMG_SYM(Send); FKMG_SYM(Recv); FKMG_SYM(AllReduce); FKMG_SYM(AllGather);
#undef FKMG_SYM
      if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.GroupStart || !a.GroupEnd || !a.Send || !a.Recv || !a.AllReduce || !a.AllGather)
        { dlclose(a.lib); a.lib = nullptr; }
    });
  return a.lib ? &a : nullptr;
}

/* contiguous ranges of `n` bins for `world` ranks: the n-th cut falls after the first bin whose running sum reaches n/world of
   the total (MSDsort.c:330-352; the same rule as fastk_b200/multigpu.py::splitters_from_hist)                              */
static void splitters(const u64 *h, int n, int world, std::vector<int> &beg)
{ std::vector<u64> cs((size_t) n);
  u64 run = 0;
  for (int i = 0; i < n; i++) { run += h[i]; cs[i] = run; }
  const u64 total = run;
  beg.assign(1,0);
  int prev = -1;
  for (int q = 1; q < world; q++)
    { const u64 target = (u64) (((unsigned __int128) total * (unsigned) q) / (unsigned) world);
      int x = (int) (std::lower_bound(cs.begin(),cs.end(),target) - cs.begin());
      x = std::max(x,prev + 1);
      if (x >= n) break;
      beg.push_back(x + 1);
      prev = x;
    }
  while ((int) beg.size() < world) beg.push_back(n);
  beg.push_back(n);
}

}  // namespace fkmg

struct MultiState
  { fkmg::Comm comm = nullptr;
    int nranks = 1, rank = 0;
    cudaEvent_t ev_rec = nullptr, ev_pay = nullptr;   /* records exchanged / base strings exchanged */
    DevBuf small;                      /* device scratch for the little collectives */
    DevBuf payload, rrec, rscr, rpay, epart, erecv;
    std::vector<int64_t> table_sizes;  /* result: table records of every rank, in rank (= key) order */
    int64_t global_ntable = 0, sent_records = 0, sent_entries = 0;
  };

#define NC(call) do { int r_ = (call); if (r_ != 0) \
    return set_err(FKGPU_E_CUDA,"%s failed at %s:%d: %s",#call,__FILE__,__LINE__,na->GetErrorString ? na->GetErrorString(r_) : "NCCL error"); } while (0)
