// Small systems through the host-pointer API: a resident "mailbox" kernel instead of a launch per call.
//
// Config 1 of BASELINE.json (solar system, N = 9, IAS15) makes ~20 million acceleration() calls per 1000 yr, each with 36
// doubles of input and 27 of output; the CPU reference needs ~0.5 us per call.  Round 1 answered every call with one kernel
// launch and one stream synchronisation from mapped pinned staging: 13.6 us per call in the real IAS15 loop.  Launch +
// synchronise is a fixed cost no kernel change removes, so this path takes the launch out of the call:
//
//   * one CTA stays resident and polls a request area in mapped pinned HOST memory;
//   * the host copies n, G, eps^2, x, m into the request area and spins on the response area;
//   * the CTA evaluates all pairs from shared memory with the direct sum's arithmetic and writes a[] to the response area.
//
// A call then costs PCIe round trips, so the protocol is built to need as few as possible -- ONE read round trip and one
// posted write, no fences on either side:
//   request   32-byte sectors of 3 doubles + a tag (the call number).  The host writes the tag of a sector last (x86 keeps
//             the store order) and the device fetches host memory in 32-byte sectors, each served from one snapshot of its
//             cache line, so a sector whose tag matches carries its payload.  Warp 0 polls the first 16 sectors with one
//             coalesced 512-byte read: enough for n <= 11 (config 1); larger systems fetch the remaining sectors in a
//             second round trip.
//   response  16-byte items (value, tag), each written with one 16-byte store; the host accepts an item when its tag is the
//             call number (it reads the tag first), so the device needs no fence between the data and a completion word.
//
// A first version used a request word, a payload read behind a system fence, a system fence after the results and a
// completion word: five round trips, 13.6 us per call in the IAS15 loop -- no better than the launch it replaced.
//
// The kernel leaves by itself after MAILBOX_IDLE_US without a request (so nothing in the process can wait on it for longer:
// cudaFree / cudaDeviceSynchronize of the application or of this library's other paths), when the host raises `quit`
// (context destruction), or when its lifetime budget runs out; the next small call starts it again.
// GRAV_B200_SMALL_MAILBOX=0 restores the launch-per-call path (direct_sum_small_host, direct_sum.cu).
#include <time.h>

#include "internal.cuh"

namespace gb {

constexpr int MB_MAX = 256;                  // particles (one target per thread)
constexpr unsigned long long MAILBOX_IDLE_US = 300;        // leave after this long without a request
constexpr unsigned long long MAILBOX_LIFE_US = 2000000;    // and after this long in any case (bounds a forgotten kernel)
constexpr int MB_HEAD = 3;                   // n, G, eps^2 precede x[3n], m[n] in the payload
constexpr int MB_SECS = (MB_HEAD + 4 * MB_MAX + 2) / 3;   // 343 request sectors for n = 256
constexpr int MB_POLL_SECS = 16;             // sectors fetched by every poll (one warp, 16 bytes per lane)

struct ReqSec { double d[3]; unsigned long long tag; };
struct RespItem { double v; unsigned long long tag; };
static_assert(sizeof(ReqSec) == 32 && sizeof(RespItem) == 16, "mailbox layout");

struct Mailbox {
    ReqSec req[MB_SECS];                     // host -> device
    RespItem resp[3 * MB_MAX];               // device -> host
    volatile int quit;                       // host -> device
};

__device__ __forceinline__ void ld_sys_v2(const void *p, unsigned long long &a, unsigned long long &b)
{
    asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void st_sys_v2(void *p, unsigned long long a, unsigned long long b)
{
    asm volatile("st.volatile.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// one ordered interaction, checked for the self term (same arithmetic as direct_sum.cu's interaction<true>)
__device__ __forceinline__ void mb_interaction(const double4 pj, bool masked, double xi, double yi, double zi, double eps2,
                                               double &ax, double &ay, double &az)
{
    const double dx = pj.x - xi, dy = pj.y - yi, dz = pj.z - zi;
    double r2 = fma(dx, dx, eps2);
    r2 = fma(dy, dy, r2);
    r2 = fma(dz, dz, r2);
    double s = inv_r3_times_m(r2, pj.w);
    if (masked) s = 0.0;
    ax = fma(s, dx, ax);
    ay = fma(s, dy, ay);
    az = fma(s, dz, az);
}

// chunk c (16 bytes) of the request area: the first half of sector c / 2 holds payload doubles 3 (c / 2) + 0, 1, the second
// half holds double 3 (c / 2) + 2 and the tag
__device__ __forceinline__ void stash_chunk(double *sP, int c, unsigned long long lo, unsigned long long hi)
{
    const int base = 3 * (c >> 1) + 2 * (c & 1);
    sP[base] = __longlong_as_double((long long)lo);
    if ((c & 1) == 0) sP[base + 1] = __longlong_as_double((long long)hi);
}

__global__ void __launch_bounds__(MB_MAX) mailbox_kernel(Mailbox *mb, unsigned long long last)
{
    __shared__ double sP[3 * MB_SECS];       // the payload, de-interleaved from the sectors
    __shared__ double4 src[MB_MAX];
    __shared__ unsigned long long s_req;
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned long long born = global_timer_ns();
    for (;;) {
        if (tid < 32) {                      // warp 0 polls: one coalesced 512-byte read of the first 16 sectors per trip
            const unsigned long long idle0 = global_timer_ns();
            unsigned long long r = ~0ull;
            for (;;) {
                unsigned long long lo, hi;
                ld_sys_v2((const char *)mb->req + 16 * lane, lo, hi);
                const unsigned long long tag0 = __shfl_sync(0xffffffffu, hi, 1);
                if (tag0 != last && tag0 != 0) {
                    const int n = (int)__longlong_as_double((long long)__shfl_sync(0xffffffffu, lo, 0));
                    const int nsec = (MB_HEAD + 4 * n + 2) / 3;
                    const bool is_tag = (lane & 1) == 1 && (lane >> 1) < min(nsec, MB_POLL_SECS);
                    if (__all_sync(0xffffffffu, !is_tag || hi == tag0)) {     // every polled sector of this request has arrived
                        stash_chunk(sP, lane, lo, hi);
                        r = tag0;
                        break;
                    }
                }
                const unsigned long long now = global_timer_ns();
                if (*(volatile int *)&mb->quit || now - idle0 > MAILBOX_IDLE_US * 1000ull || now - born > MAILBOX_LIFE_US * 1000ull) break;
            }
            if (lane == 0) s_req = r;
        }
        __syncthreads();
        const unsigned long long r = s_req;
        if (r == ~0ull) break;
        const int n = (int)sP[0];
        const int nsec = (MB_HEAD + 4 * n + 2) / 3;
        for (int S = MB_POLL_SECS + tid; S < nsec; S += MB_MAX) {       // larger systems: the remaining sectors, one per thread
            unsigned long long q[4];
            do {
                ld_sys_v2((const char *)&mb->req[S], q[0], q[1]);
                ld_sys_v2((const char *)&mb->req[S] + 16, q[2], q[3]);
            } while (q[3] != r);
            stash_chunk(sP, 2 * S, q[0], q[1]);
            stash_chunk(sP, 2 * S + 1, q[2], q[3]);
        }
        __syncthreads();
        const double G = sP[1], eps2 = sP[2];
        if (tid < n) src[tid] = make_double4(sP[MB_HEAD + 3 * tid], sP[MB_HEAD + 3 * tid + 1], sP[MB_HEAD + 3 * tid + 2], sP[MB_HEAD + 3 * n + tid]);
        __syncthreads();
        if (tid < n) {
            const double4 me = src[tid];
            double ax = 0.0, ay = 0.0, az = 0.0;
            for (int j = 0; j < n; j++) mb_interaction(src[j], j == tid, me.x, me.y, me.z, eps2, ax, ay, az);
            st_sys_v2(&mb->resp[3 * tid + 0], (unsigned long long)__double_as_longlong(G * ax), r);
            st_sys_v2(&mb->resp[3 * tid + 1], (unsigned long long)__double_as_longlong(G * ay), r);
            st_sys_v2(&mb->resp[3 * tid + 2], (unsigned long long)__double_as_longlong(G * az), r);
        }
        last = r;
        __syncthreads();                     // sP / src are rewritten by the next request
    }
}

struct MailboxState {
    Mailbox *mb = nullptr;          // mapped pinned host memory (UVA: the same pointer on the device)
    cudaStream_t stream = nullptr;
    unsigned long long seq = 0;
    bool launched = false;          // a kernel was launched and has not been seen finished yet
};

static double now_s()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static int mailbox_start(grav_b200_ctx *c, MailboxState *s, unsigned long long last_done)
{
    s->mb->quit = 0;
    __sync_synchronize();
    mailbox_kernel<<<1, MB_MAX, 0, s->stream>>>(s->mb, last_done);
    GB_LAUNCH_CHECK();
    count_launch();
    s->launched = true;
    (void)c;
    return GRAV_B200_OK;
}

// Stops the resident kernel (if any) and waits for it; called before any other work of the context and on destruction.
int mailbox_stop(grav_b200_ctx *c)
{
    MailboxState *s = (MailboxState *)c->mailbox;
    if (!s || !s->launched) return GRAV_B200_OK;
    s->mb->quit = 1;
    __sync_synchronize();
    GB_CUDA(cudaStreamSynchronize(s->stream));
    s->launched = false;
    return GRAV_B200_OK;
}

void mailbox_free(grav_b200_ctx *c)
{
    MailboxState *s = (MailboxState *)c->mailbox;
    if (!s) return;
    mailbox_stop(c);
    if (s->stream) cudaStreamDestroy(s->stream);
    if (s->mb) cudaFreeHost(s->mb);
    delete s;
    c->mailbox = nullptr;
}

// a, x, m: the caller's host arrays; n <= 256.
int mailbox_pairwise(grav_b200_ctx *c, double *a, int n, const double *x, const double *m, double G, double eps)
{
    if (n > MB_MAX) { set_error("mailbox path: n = %d > %d", n, MB_MAX); return GRAV_B200_EINVAL; }
    MailboxState *s = (MailboxState *)c->mailbox;
    if (!s) {
        s = new MailboxState();
        cudaError_t e = cudaHostAlloc((void **)&s->mb, sizeof(Mailbox), cudaHostAllocMapped);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            if (s->mb) cudaFreeHost(s->mb);
            delete s;
            return cuda_fail(e, "mailbox setup", __FILE__, __LINE__);
        }
        memset((void *)s->mb, 0, sizeof(Mailbox));
        c->mailbox = s;
    }
    Mailbox *mb = s->mb;
    const unsigned long long seq = ++s->seq;
    // payload stream: n, G, eps^2, x[3n], m[n]; three doubles per 32-byte sector, the sector's tag written last
    const int total = MB_HEAD + 4 * n;
    auto payload = [&](int k) -> double {
        if (k == 0) return (double)n;
        if (k == 1) return G;
        if (k == 2) return eps * eps;
        k -= MB_HEAD;
        return k < 3 * n ? x[k] : m[k - 3 * n];
    };
    const int nsec = (total + 2) / 3;
    for (int L = nsec - 1; L >= 0; L--) {          // sector 0 (the one the poll looks at first) last
        ReqSec *ln = &mb->req[L];
        for (int k = 0; k < 3; k++) ln->d[k] = (3 * L + k < total) ? payload(3 * L + k) : 0.0;
        __asm__ __volatile__("" ::: "memory");     // x86 keeps the store order; the compiler must too
        *(volatile unsigned long long *)&ln->tag = seq;
    }
    __sync_synchronize();
    if (!s->launched) GB_TRY(mailbox_start(c, s, seq - 1));
    // collect the response items; now and then make sure the kernel is still there (it leaves when idle)
    const double t0 = now_s();
    unsigned spins = 0;
    for (int k = 0; k < 3 * n;) {
        const volatile RespItem *it = &mb->resp[k];
        if (it->tag == seq) {
            __asm__ __volatile__("" ::: "memory");
            a[k] = it->v;
            k++;
            continue;
        }
        if ((++spins & 0x3ff) == 0) {
            if (cudaStreamQuery(s->stream) == cudaSuccess) {      // the kernel left (idle / lifetime) without seeing this request
                s->launched = false;
                GB_TRY(mailbox_start(c, s, seq - 1));
            } else {
                cudaGetLastError();                                 // cudaErrorNotReady is the normal answer
            }
            if (now_s() - t0 > 5.0) {
                mb->quit = 1;
                set_error("small-system mailbox kernel did not answer within 5 s");
                return GRAV_B200_ECUDA;
            }
        }
        __builtin_ia32_pause();
    }
    return GRAV_B200_OK;
}

}  // namespace gb
