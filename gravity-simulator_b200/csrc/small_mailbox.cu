// Small systems through the host-pointer API: a resident "mailbox" kernel instead of a launch per call.
//
// Config 1 of BASELINE.json (solar system, N = 9, IAS15) makes ~8 million acceleration() calls, each with 36 doubles of
// input and 27 of output; the CPU reference needs ~2 us per call.  Round 1 answered every call with one kernel launch and
// one stream synchronisation from mapped pinned staging: ~15 us, i.e. config 1 ran 28x slower than the reference through the
// drop-in.  Launch + synchronise is a fixed cost no kernel change removes, so this path takes the launch out of the call:
//
//   * one CTA stays resident and polls a request word in mapped pinned HOST memory;
//   * the host copies x, m into the mailbox, bumps the request word, and spins on the completion word;
//   * the CTA reads the request (two PCIe round trips: the word, then the payload), evaluates all pairs from shared memory
//     with the direct sum's arithmetic, writes a[] to the mailbox, fences and bumps the completion word.
//
// The kernel leaves by itself after MAILBOX_IDLE_US without a request (so nothing in the process can wait on it for longer:
// cudaFree / cudaDeviceSynchronize of the application or of this library's other paths), when the host raises `quit` (any
// larger call on the same context, context destruction), or when its lifetime budget runs out; the next small call starts it
// again.  GRAV_B200_SMALL_MAILBOX=0 restores the launch-per-call path (direct_sum_small_host, direct_sum.cu).
#include <time.h>

#include "internal.cuh"

namespace gb {

constexpr int MB_MAX = 256;                  // particles (one target per thread)
constexpr unsigned long long MAILBOX_IDLE_US = 300;        // leave after this long without a request
constexpr unsigned long long MAILBOX_LIFE_US = 2000000;    // and after this long in any case (bounds a forgotten kernel)

struct Mailbox {
    // host -> device
    volatile unsigned long long req;         // request number, written LAST by the host
    int n;
    int quit;
    double G, eps2;
    double x[3 * MB_MAX];
    double m[MB_MAX];
    // device -> host
    double a[3 * MB_MAX];
    volatile unsigned long long ack;         // number of the last completed request, written LAST by the device
    volatile unsigned long long alive;       // set by the kernel when it starts polling, cleared when it leaves
};

__device__ __forceinline__ unsigned long long ld_sys_u64(const volatile unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
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

__global__ void __launch_bounds__(MB_MAX) mailbox_kernel(Mailbox *mb)
{
    __shared__ double4 src[MB_MAX];
    __shared__ unsigned long long s_req;
    __shared__ int s_n;
    __shared__ double s_G, s_eps2;
    const int tid = threadIdx.x;
    unsigned long long last = 0;
    const unsigned long long born = global_timer_ns();
    if (tid == 0) {
        last = ld_sys_u64(&mb->ack);
        mb->alive = 1;
        __threadfence_system();
    }
    for (;;) {
        if (tid == 0) {
            const unsigned long long idle0 = global_timer_ns();
            unsigned long long r;
            for (;;) {
                r = ld_sys_u64(&mb->req);
                if (r != last) break;
                const unsigned long long now = global_timer_ns();
                if (*(volatile int *)&mb->quit || now - idle0 > MAILBOX_IDLE_US * 1000ull || now - born > MAILBOX_LIFE_US * 1000ull) { r = ~0ull; break; }
            }
            if (r != ~0ull) {
                __threadfence_system();        // the payload was written before the request word
                s_n = *(volatile int *)&mb->n;
                s_G = *(volatile double *)&mb->G;
                s_eps2 = *(volatile double *)&mb->eps2;
            }
            s_req = r;
        }
        __syncthreads();
        const unsigned long long r = s_req;
        if (r == ~0ull) break;
        const int n = s_n;
        if (tid < n) {
            const volatile double *hx = mb->x, *hm = mb->m;
            src[tid] = make_double4(hx[3 * tid], hx[3 * tid + 1], hx[3 * tid + 2], hm[tid]);
        }
        __syncthreads();
        if (tid < n) {
            const double4 me = src[tid];
            double ax = 0.0, ay = 0.0, az = 0.0;
            for (int j = 0; j < n; j++) mb_interaction(src[j], j == tid, me.x, me.y, me.z, s_eps2, ax, ay, az);
            volatile double *ha = mb->a;
            ha[3 * tid + 0] = s_G * ax;
            ha[3 * tid + 1] = s_G * ay;
            ha[3 * tid + 2] = s_G * az;
            __threadfence_system();            // results are in host memory before the completion word
        }
        __syncthreads();
        if (tid == 0) {
            mb->ack = r;
            __threadfence_system();
            last = r;
        }
    }
    if (tid == 0) {
        mb->alive = 0;
        __threadfence_system();
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

static int mailbox_start(grav_b200_ctx *c, MailboxState *s)
{
    s->mb->quit = 0;
    __sync_synchronize();
    mailbox_kernel<<<1, MB_MAX, 0, s->stream>>>(s->mb);
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
    memcpy(mb->x, x, sizeof(double) * 3 * (size_t)n);
    memcpy(mb->m, m, sizeof(double) * (size_t)n);
    mb->n = n;
    mb->G = G;
    mb->eps2 = eps * eps;
    const unsigned long long seq = ++s->seq;
    __sync_synchronize();                  // payload before the request word (x86: a compiler barrier + store ordering)
    mb->req = seq;
    __sync_synchronize();
    if (!s->launched) GB_TRY(mailbox_start(c, s));
    // wait for the completion word; now and then make sure the kernel is still there (it leaves when idle)
    const double t0 = now_s();
    unsigned spins = 0;
    while (mb->ack != seq) {
        if ((++spins & 0x3ff) == 0) {
            if (cudaStreamQuery(s->stream) == cudaSuccess) {      // the kernel left (idle / lifetime) without seeing this request
                s->launched = false;
                if (mb->ack == seq) break;
                GB_TRY(mailbox_start(c, s));
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
    __sync_synchronize();
    memcpy(a, mb->a, sizeof(double) * 3 * (size_t)n);
    return GRAV_B200_OK;
}

}  // namespace gb
