/*  numa_probe.cu -- where does the device->host copy of the table go when several GPUs fetch at once?
 *  One thread per GPU copies BYTES from device memory into a page-locked host buffer; the buffer is allocated
 *    mode 0: as the process happens to run (what the library did),
 *    mode 1: under set_mempolicy(MPOL_BIND, node of the GPU)  (may be refused in a container),
 *    mode 2: by a thread pinned to the CPUs of the GPU's node (first touch, then cudaHostRegister).
 *  Prints per-GPU times and the aggregate for: each GPU alone, some pairs, all GPUs.
 *  Build: nvcc -O2 -o numa_probe tools/numa_probe.cu -lpthread                                                  */
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <atomic>
#include <chrono>
#include <sched.h>
#include <unistd.h>
#include <sys/syscall.h>
#include <sys/mman.h>

static const size_t BYTES = (size_t) 1400 << 20;

static int gpu_node(int dev)
{ char bus[64];
  if (cudaDeviceGetPCIBusId(bus,sizeof(bus),dev) != cudaSuccess) return -1;
  for (char *p = bus; *p; p++) if (*p >= 'A' && *p <= 'Z') *p += 32;
  std::string path = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
  FILE *f = fopen(path.c_str(),"r");
  if (!f) return -1;
  int n = -1;
  if (fscanf(f,"%d",&n) != 1) n = -1;
  fclose(f);
  return n;
}

static bool node_cpus(int node, cpu_set_t *set)
{ char path[128];
  snprintf(path,sizeof(path),"/sys/devices/system/node/node%d/cpulist",node);
  FILE *f = fopen(path,"r");
  if (!f) return false;
  char buf[4096];
  if (!fgets(buf,sizeof(buf),f)) { fclose(f); return false; }
  fclose(f);
  CPU_ZERO(set);
  for (char *tok = strtok(buf,",\n"); tok; tok = strtok(NULL,",\n"))
    { int a, b;
      if (sscanf(tok,"%d-%d",&a,&b) == 2) { for (int c = a; c <= b; c++) CPU_SET(c,set); }
      else if (sscanf(tok,"%d",&a) == 1) CPU_SET(a,set);
    }
  return true;
}

static long set_policy(int mode, int node)   /* mode 2 = MPOL_BIND, 0 = default */
{ unsigned long mask[16] = {0};
  if (mode != 0 && node >= 0) mask[node / 64] |= 1ul << (node % 64);
  return syscall(SYS_set_mempolicy,mode,mode ? mask : NULL,mode ? 1024 : 0);
}

struct Gpu { int dev, node; void *d; void *h[3]; cudaStream_t s; cudaEvent_t a, b; };

static std::atomic<int> arrive{0};
static void barrier(int n, int phase) { arrive.fetch_add(1); while (arrive.load() < n * phase) std::this_thread::yield(); }

int main()
{ int n = 0;
  cudaGetDeviceCount(&n);
  printf("gpus %d\n",n);
  std::vector<Gpu> g(n);
  for (int i = 0; i < n; i++) { g[i].dev = i; g[i].node = gpu_node(i); printf("gpu %d numa_node %d\n",i,g[i].node); }
  cpu_set_t cur; sched_getaffinity(0,sizeof(cur),&cur);
  printf("cpus allowed %d\n",CPU_COUNT(&cur));
  long rp = set_policy(2,g[0].node >= 0 ? g[0].node : 0);
  printf("set_mempolicy(MPOL_BIND) -> %ld (%s)\n",rp,rp ? strerror(errno) : "ok");
  set_policy(0,0);

  /* allocate */
  std::vector<std::thread> th;
  for (int i = 0; i < n; i++) th.emplace_back([&,i]()
    { Gpu &x = g[i];
      cudaSetDevice(x.dev);
      cudaMalloc(&x.d,BYTES); cudaMemset(x.d,i + 1,BYTES);
      cudaStreamCreate(&x.s); cudaEventCreate(&x.a); cudaEventCreate(&x.b);
      cudaMallocHost(&x.h[0],BYTES); memset(x.h[0],0,BYTES);
      x.h[1] = NULL;
      if (x.node >= 0 && set_policy(2,x.node) == 0)
        { cudaMallocHost(&x.h[1],BYTES); memset(x.h[1],0,BYTES); set_policy(0,0); }
      x.h[2] = NULL;
      cpu_set_t set, old;
      sched_getaffinity(0,sizeof(old),&old);
      if (x.node >= 0 && node_cpus(x.node,&set) && sched_setaffinity(0,sizeof(set),&set) == 0)
        { void *p = mmap(NULL,BYTES,PROT_READ | PROT_WRITE,MAP_PRIVATE | MAP_ANONYMOUS,-1,0);
          if (p != MAP_FAILED)
            { memset(p,0,BYTES);
              if (cudaHostRegister(p,BYTES,cudaHostRegisterDefault) == cudaSuccess) x.h[2] = p; else cudaGetLastError();
            }
          sched_setaffinity(0,sizeof(old),&old);
        }
      cudaDeviceSynchronize();
    });
  for (auto &t : th) t.join();
  th.clear();
  for (int i = 0; i < n; i++) printf("gpu %d buffers: default %p bind %p firsttouch %p\n",i,g[i].h[0],g[i].h[1],g[i].h[2]);

  auto run = [&](std::vector<int> set, int mode, const char *name)
    { std::vector<float> ms(n,0.f);
      arrive = 0;
      int m = (int) set.size();
      std::vector<std::thread> t2;
      auto t0 = std::chrono::steady_clock::now();
      for (int i : set) t2.emplace_back([&,i]()
        { Gpu &x = g[i];
          cudaSetDevice(x.dev);
          if (x.h[mode] == NULL) return;
          for (int rep = 0; rep < 3; rep++)
            { barrier(m,rep + 1);
              cudaEventRecord(x.a,x.s);
              cudaMemcpyAsync(x.h[mode],x.d,BYTES,cudaMemcpyDeviceToHost,x.s);
              cudaEventRecord(x.b,x.s);
              cudaStreamSynchronize(x.s);
              float e; cudaEventElapsedTime(&e,x.a,x.b);
              ms[i] = e;
            }
        });
      for (auto &t : t2) t.join();
      (void) t0;
      float mx = 0;
      printf("%-12s mode %d:",name,mode);
      for (int i : set) { printf(" g%d %.1f",i,ms[i]); if (ms[i] > mx) mx = ms[i]; }
      if (mx > 0) printf("  | max %.1f ms, aggregate %.1f GB/s\n",mx,m * (BYTES / 1e6) / mx); else printf("  | n/a\n");
    };

  for (int mode = 0; mode < 3; mode++)
    { run({0},mode,"alone0");
      if (n > 1) run({n - 1},mode,"aloneLast");
      if (n > 1) run({0,1},mode,"pair01");
      if (n > 2) run({0,2},mode,"pair02");
      if (n > 4) run({0,4},mode,"pair04");
      if (n > 3) run({0,1,2,3},mode,"quad0123");
      if (n > 4) { std::vector<int> all; for (int i = 0; i < n; i++) all.push_back(i); run(all,mode,"all"); }
    }
  return 0;
}
