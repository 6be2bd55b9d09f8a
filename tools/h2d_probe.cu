// h2d_probe — host-to-device copy ceiling of this box, the way bench.py's e2e leg uses it: N processes (one per GPU),
// each streaming frames from its own page-locked buffer with cudaMemcpyAsync, all started together.
//
//   h2d_probe [--gpus N] [--mb 2048] [--chunk-mb 332] [--reps 3] [--modes default,local,node0,node1,interleave]
//             [--streams 1,2] [--first-gpu G]
//
// Placement modes of the page-locked buffer:
//   default     cudaHostAlloc (first touch by the CUDA driver; whatever the process's memory policy gives)
//   local       mmap + mbind(MPOL_BIND, NUMA node of the GPU's PCI device) + first touch + cudaHostRegister
//   nodeK       the same, bound to node K
//   interleave  mbind(MPOL_INTERLEAVE) over all nodes the process may allocate from
// Prints one JSON line: the topology it saw (GPU -> PCI bus id -> NUMA node, Mems_allowed, Cpus_allowed) and per
// (mode, streams): per-GPU GB/s, the aggregate, and where the pages really landed (move_pages query on a sample).
// No kernel is launched; nothing waits on another process on the device (start-up is a host-side barrier in shared memory).
#include <cuda_runtime.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <sys/wait.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#ifndef MPOL_BIND
#define MPOL_BIND 2
#define MPOL_INTERLEAVE 3
#endif
#ifndef MPOL_MF_STRICT
#define MPOL_MF_STRICT 1
#endif

static long sys_mbind(void *addr, unsigned long len, int mode, const unsigned long *mask, unsigned long maxnode, unsigned flags) {
    return syscall(SYS_mbind, addr, len, mode, mask, maxnode, flags);
}
static long sys_move_pages(int pid, unsigned long count, void **pages, const int *nodes, int *status, int flags) {
    return syscall(SYS_move_pages, pid, count, pages, nodes, status, flags);
}

static std::string read_first_line(const std::string &path) {
    std::ifstream f(path);
    std::string s;
    std::getline(f, s);
    return s;
}
static std::string status_field(const char *key) {
    std::ifstream f("/proc/self/status");
    std::string line;
    const size_t kl = strlen(key);
    while (std::getline(f, line))
        if (!line.compare(0, kl, key)) {
            size_t p = line.find_first_not_of(" \t", kl + 1);
            return p == std::string::npos ? "" : line.substr(p);
        }
    return "";
}
// "0-1,4" -> bit mask (nodes < 64 are enough here)
static unsigned long parse_list_mask(const std::string &s) {
    unsigned long m = 0;
    std::stringstream ss(s);
    std::string tok;
    while (std::getline(ss, tok, ',')) {
        int a, b;
        if (sscanf(tok.c_str(), "%d-%d", &a, &b) == 2) {
            for (int i = a; i <= b && i < 64; ++i) m |= 1ul << i;
        } else if (sscanf(tok.c_str(), "%d", &a) == 1 && a < 64)
            m |= 1ul << a;
    }
    return m;
}

struct Shared {
    std::atomic<int> arrived[64];
    std::atomic<int> go[64];
    double gbs[16];
    int node_hist[16][8];
    int err[16];
    char busid[16][32];
    int numa[16];
};

static void barrier(Shared *S, int phase, int n) {
    S->arrived[phase].fetch_add(1);
    while (S->arrived[phase].load() < n) sched_yield();
}

struct Buf {
    void *p = nullptr;
    size_t bytes = 0;
    bool registered = false;
};
static int alloc_buf(Buf &b, size_t bytes, const std::string &mode, int gpu_node, unsigned long allowed) {
    b.bytes = bytes;
    if (mode == "default") return cudaHostAlloc(&b.p, bytes, cudaHostAllocPortable) == cudaSuccess ? 0 : 1;
    void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) return 2;
    unsigned long mask = 0;
    int pol = MPOL_BIND;
    if (mode == "local")
        mask = gpu_node >= 0 ? 1ul << gpu_node : 0;
    else if (mode == "interleave") {
        mask = allowed;
        pol = MPOL_INTERLEAVE;
    } else if (!mode.compare(0, 4, "node"))
        mask = 1ul << atoi(mode.c_str() + 4);
    if (mask && sys_mbind(p, bytes, pol, &mask, 64, 0) != 0) {
        munmap(p, bytes);
        return 3;  // the cpuset does not allow that node (or the kernel refuses)
    }
    memset(p, 1, bytes);  // first touch under the policy
    if (cudaHostRegister(p, bytes, cudaHostRegisterPortable) != cudaSuccess) {
        munmap(p, bytes);
        return 4;
    }
    b.p = p;
    b.registered = true;
    return 0;
}
static void free_buf(Buf &b) {
    if (!b.p) return;
    if (b.registered) {
        cudaHostUnregister(b.p);
        munmap(b.p, b.bytes);
    } else
        cudaFreeHost(b.p);
    b = Buf{};
}

int main(int argc, char **argv) {
    int ngpu = 1, reps = 3, first_gpu = 0;
    size_t mb = 2048, chunk_mb = 332;
    std::string modes = "default,local,interleave", streams_s = "1,2";
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() { return i + 1 < argc ? std::string(argv[++i]) : std::string(); };
        if (a == "--gpus") ngpu = atoi(next().c_str());
        else if (a == "--mb") mb = strtoull(next().c_str(), nullptr, 10);
        else if (a == "--chunk-mb") chunk_mb = strtoull(next().c_str(), nullptr, 10);
        else if (a == "--reps") reps = atoi(next().c_str());
        else if (a == "--modes") modes = next();
        else if (a == "--streams") streams_s = next();
        else if (a == "--first-gpu") first_gpu = atoi(next().c_str());
    }
    if (ngpu < 1 || ngpu > 16) return 2;
    std::vector<std::string> mode_list;
    {
        std::stringstream ss(modes);
        std::string t;
        while (std::getline(ss, t, ',')) mode_list.push_back(t);
    }
    std::vector<int> stream_list;
    {
        std::stringstream ss(streams_s);
        std::string t;
        while (std::getline(ss, t, ',')) stream_list.push_back(std::max(1, atoi(t.c_str())));
    }
    const unsigned long allowed = parse_list_mask(status_field("Mems_allowed_list"));
    const int ncombo = (int)(mode_list.size() * stream_list.size());
    if (2 * ncombo + 2 > 64) return 2;
    Shared *S = (Shared *)mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    memset((void *)S, 0, sizeof(Shared));
    // results[combo][gpu]
    double *all_gbs = (double *)mmap(nullptr, sizeof(double) * 64 * 16, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    int *all_err = (int *)mmap(nullptr, sizeof(int) * 64 * 16, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    int *all_hist = (int *)mmap(nullptr, sizeof(int) * 64 * 16 * 8, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    double *all_wall = (double *)mmap(nullptr, sizeof(double) * 64 * 16, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    memset(all_hist, 0, sizeof(int) * 64 * 16 * 8);

    std::vector<pid_t> kids;
    for (int g = 0; g < ngpu; ++g) {
        pid_t pid = fork();
        if (pid == 0) {
            const int dev = first_gpu + g;
            if (cudaSetDevice(dev) != cudaSuccess) {
                S->err[g] = 100;
                for (int ph = 0; ph < 2 * ncombo + 1; ++ph) barrier(S, ph, ngpu);
                _exit(0);
            }
            char bus[32] = {0};
            cudaDeviceGetPCIBusId(bus, sizeof bus, dev);
            for (char *c = bus; *c; ++c) *c = (char)tolower(*c);
            strncpy(S->busid[g], bus, 31);
            int node = -1;
            {
                std::string s = read_first_line(std::string("/sys/bus/pci/devices/") + bus + "/numa_node");
                if (!s.empty()) node = atoi(s.c_str());
            }
            S->numa[g] = node;
            const size_t bytes = mb << 20, chunk = std::min(bytes, chunk_mb << 20);
            void *d[2] = {nullptr, nullptr};
            cudaMalloc(&d[0], chunk);
            cudaMalloc(&d[1], chunk);
            cudaStream_t st[4];
            for (auto &s : st) cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
            int combo = 0;
            barrier(S, 0, ngpu);
            for (const std::string &mode : mode_list) {
                Buf b;
                const int rc = alloc_buf(b, bytes, mode, node, allowed);
                if (rc == 0) {  // where did the pages land?
                    const int ns = 64;
                    void *pages[ns];
                    int status[ns];
                    for (int q = 0; q < ns; ++q) pages[q] = (char *)b.p + (bytes / ns) * q;
                    if (sys_move_pages(0, ns, pages, nullptr, status, 0) == 0)
                        for (int q = 0; q < ns; ++q)
                            if (status[q] >= 0 && status[q] < 8)
                                for (size_t si = 0; si < stream_list.size(); ++si) all_hist[((combo + si) * 16 + g) * 8 + status[q]]++;
                }
                for (int nst : stream_list) {
                    all_err[combo * 16 + g] = rc;
                    barrier(S, 1 + 2 * combo, ngpu);
                    double gbs = 0;
                    if (rc == 0) {
                        // warm-up
                        cudaMemcpyAsync(d[0], b.p, chunk, cudaMemcpyHostToDevice, st[0]);
                        cudaStreamSynchronize(st[0]);
                    }
                    barrier(S, 2 + 2 * combo, ngpu);
                    auto t0 = std::chrono::steady_clock::now();
                    if (rc == 0) {
                        size_t moved = 0;
                        int k = 0;
                        for (int r = 0; r < reps; ++r)
                            for (size_t off = 0; off + chunk <= bytes; off += chunk, ++k) {
                                cudaMemcpyAsync(d[k % 2], (char *)b.p + off, chunk, cudaMemcpyHostToDevice, st[k % nst]);
                                moved += chunk;
                            }
                        for (int q = 0; q < nst; ++q) cudaStreamSynchronize(st[q]);
                        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                        gbs = moved / dt / 1e9;
                    }
                    all_gbs[combo * 16 + g] = gbs;
                    all_wall[combo * 16 + g] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                    ++combo;
                }
                free_buf(b);
            }
            _exit(0);
        }
        kids.push_back(pid);
    }
    for (pid_t p : kids) {
        int st = 0;
        waitpid(p, &st, 0);
    }
    // ---- report ------------------------------------------------------------------------------------------------
    printf("{\"tool\": \"h2d_probe\", \"gpus\": %d, \"mb_per_gpu\": %zu, \"chunk_mb\": %zu, \"reps\": %d, ", ngpu, mb, chunk_mb, reps);
    printf("\"mems_allowed\": \"%s\", \"cpus_allowed\": \"%s\", ", status_field("Mems_allowed_list").c_str(), status_field("Cpus_allowed_list").c_str());
    printf("\"numa_nodes_online\": \"%s\", ", read_first_line("/sys/devices/system/node/online").c_str());
    printf("\"gpu\": [");
    for (int g = 0; g < ngpu; ++g) printf("%s{\"bus\": \"%s\", \"numa_node\": %d}", g ? ", " : "", S->busid[g], S->numa[g]);
    printf("], \"runs\": [");
    int combo = 0;
    for (size_t mi = 0; mi < mode_list.size(); ++mi)
        for (size_t si = 0; si < stream_list.size(); ++si, ++combo) {
            double agg = 0, mn = 1e30;
            int err = 0;
            printf("%s{\"mode\": \"%s\", \"streams\": %d, \"per_gpu_gbs\": [", combo ? ", " : "", mode_list[mi].c_str(), stream_list[si]);
            for (int g = 0; g < ngpu; ++g) {
                const double v = all_gbs[combo * 16 + g];
                printf("%s%.2f", g ? ", " : "", v);
                agg += v;
                mn = std::min(mn, v);
                err = std::max(err, all_err[combo * 16 + g]);
            }
            double wmax = 0, bytes_moved = 0;
            for (int g = 0; g < ngpu; ++g) {
                wmax = std::max(wmax, all_wall[combo * 16 + g]);
                bytes_moved += all_gbs[combo * 16 + g] * all_wall[combo * 16 + g];
            }
            printf("], \"aggregate_gbs\": %.2f, \"aggregate_gbs_by_slowest\": %.2f, \"min_gbs\": %.2f, \"alloc_error\": %d, \"pages_on_node\": [", agg,
                   wmax > 0 ? bytes_moved / wmax : 0.0, mn, err);
            for (int g = 0; g < ngpu; ++g) {
                printf("%s[", g ? ", " : "");
                for (int q = 0; q < 8; ++q) printf("%s%d", q ? "," : "", all_hist[(combo * 16 + g) * 8 + q]);
                printf("]");
            }
            printf("]}");
        }
    printf("]}\n");
    return 0;
}
