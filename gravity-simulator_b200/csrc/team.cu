// Device teams: several GPUs of one box driven from ONE calling thread through the ordinary context API.
//
// The reference's callers are single-process and single-threaded: acceleration() is called from integrator.c /
// integrator_ias15.c / integrator_rk_embedded.c (src/integrator.c:963,1021 ...) and launch_simulation_python from one
// Python worker thread (grav_sim/simulator.py:66-102).  They cannot spawn one process per GPU, so the sharded paths of
// this library (direct sum by target range with an all-gather of positions, Barnes-Hut with a replicated build and a
// sharded walk) would be out of their reach.  A team makes them reachable without touching the callers:
//
//   grav_b200_ctx_create_team(&ctx, k, devices)   returns a LEADER context (rank 0, on the calling thread's device)
//   and starts k-1 worker threads, each owning the context of one more device (rank r of k; the NCCL communicator is
//   built with ncclCommInitRank from the k threads, exactly as k processes would).  Every grav_b200_ctx_* entry called
//   on the leader is forwarded to all members: the workers execute the same call on their own context on their own
//   thread, the caller executes it on the leader, and the call returns when all have returned.  Host arrays are shared
//   memory here, so uploads read the caller's buffers from k threads and downloads write only each rank's own slice.
//
// GRAV_B200_DEVICES=k (default 1) makes the host-pointer one-shots -- i.e. the drop-in acceleration() -- and the
// resident time loops behind leapfrog()/euler()/euler_cromer()/rk4() use a team of devices 0..k-1.
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "internal.cuh"

namespace gb {

struct Team {
    int k = 0;
    std::vector<int> device;
    std::vector<grav_b200_ctx *> ctx;      // [0] = leader
    std::vector<std::thread> th;           // workers of ranks 1..k-1
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    std::function<int(grav_b200_ctx *)> fn;
    uint64_t seq = 0;
    int pending = 0;
    bool stop = false;
    std::vector<int> rc;
    std::vector<std::string> err;
    unsigned char uid[128];
};

static thread_local bool t_in_team = false;

bool team_active(const grav_b200_ctx *c) { return c && c->team && !t_in_team; }

static void worker_main(Team *T, int r)
{
    t_in_team = true;
    // the member context is created here so that the device binding, the stream and the NCCL rank belong to this thread
    int rc = grav_b200_ctx_create(&T->ctx[r], T->device[r], r, T->k, T->uid);
    if (rc == GRAV_B200_OK) T->ctx[r]->team = T;
    {
        std::lock_guard<std::mutex> lk(T->mu);
        T->rc[r] = rc;
        T->err[r] = rc ? grav_b200_last_error() : "";
        if (--T->pending == 0) T->cv_done.notify_one();
    }
    uint64_t seen = 0;
    for (;;) {
        std::function<int(grav_b200_ctx *)> f;
        {
            std::unique_lock<std::mutex> lk(T->mu);
            T->cv_go.wait(lk, [&] { return T->stop || T->seq != seen; });
            if (T->stop) break;
            seen = T->seq;
            f = T->fn;
        }
        rc = T->ctx[r] ? f(T->ctx[r]) : GRAV_B200_EINVAL;
        std::lock_guard<std::mutex> lk(T->mu);
        T->rc[r] = rc;
        T->err[r] = rc ? grav_b200_last_error() : "";
        if (--T->pending == 0) T->cv_done.notify_one();
    }
    if (T->ctx[r]) {
        T->ctx[r]->team = nullptr;
        grav_b200_ctx_destroy(T->ctx[r]);
        T->ctx[r] = nullptr;
    }
}

// Runs f on every member (the leader's on the calling thread) and returns the first failure.
int team_run(grav_b200_ctx *c, const std::function<int(grav_b200_ctx *)> &f)
{
    Team *T = c->team;
    {
        std::lock_guard<std::mutex> lk(T->mu);
        T->fn = f;
        T->seq++;
        T->pending = T->k - 1;
    }
    T->cv_go.notify_all();
    t_in_team = true;
    const int rc0 = f(T->ctx[0]);
    t_in_team = false;
    std::string err0 = rc0 ? grav_b200_last_error() : "";
    {
        std::unique_lock<std::mutex> lk(T->mu);
        T->cv_done.wait(lk, [&] { return T->pending == 0; });
    }
    if (rc0 != GRAV_B200_OK) { set_error("%s", err0.c_str()); return rc0; }
    for (int r = 1; r < T->k; r++) {
        if (T->rc[r] != GRAV_B200_OK) {
            set_error("%s (device team, rank %d on device %d)", T->err[r].c_str(), r, T->device[r]);
            return T->rc[r];
        }
    }
    return GRAV_B200_OK;
}

void team_destroy(grav_b200_ctx *leader)
{
    Team *T = leader->team;
    if (!T) return;
    {
        std::lock_guard<std::mutex> lk(T->mu);
        T->stop = true;
    }
    T->cv_go.notify_all();
    for (auto &t : T->th) if (t.joinable()) t.join();
    leader->team = nullptr;
    delete T;
}

}  // namespace gb

using namespace gb;

extern "C" int grav_b200_ctx_create_team(grav_b200_ctx **out, int num_devices, const int *devices)
{
    if (!out) { set_error("NULL out pointer"); return GRAV_B200_EINVAL; }
    *out = nullptr;
    const int ndev = grav_b200_device_count();
    if (ndev == 0) { set_error("no CUDA device available; libgrav_b200 has no CPU fallback"); return GRAV_B200_ENODEV; }
    if (num_devices < 1 || num_devices > ndev) { set_error("a team of %d devices on a box with %d", num_devices, ndev); return GRAV_B200_EINVAL; }
    std::vector<int> dev(num_devices);
    for (int r = 0; r < num_devices; r++) {
        dev[r] = devices ? devices[r] : r;
        if (dev[r] < 0 || dev[r] >= ndev) { set_error("device %d out of range [0,%d)", dev[r], ndev); return GRAV_B200_ENODEV; }
        for (int q = 0; q < r; q++) if (dev[q] == dev[r]) { set_error("device %d listed twice", dev[r]); return GRAV_B200_EINVAL; }
    }
    if (num_devices == 1) return grav_b200_ctx_create(out, dev[0], 0, 1, nullptr);
    Team *T = new Team();
    T->k = num_devices;
    T->device = dev;
    T->ctx.assign(num_devices, nullptr);
    T->rc.assign(num_devices, 0);
    T->err.assign(num_devices, "");
    int rc = grav_b200_nccl_unique_id(T->uid);
    if (rc != GRAV_B200_OK) { delete T; return rc; }
    T->pending = num_devices - 1;
    for (int r = 1; r < num_devices; r++) T->th.emplace_back(worker_main, T, r);
    rc = grav_b200_ctx_create(&T->ctx[0], dev[0], 0, num_devices, T->uid);   // joins the workers in ncclCommInitRank
    std::string err0 = rc ? grav_b200_last_error() : "";
    {
        std::unique_lock<std::mutex> lk(T->mu);
        T->cv_done.wait(lk, [&] { return T->pending == 0; });
    }
    int bad = rc;
    std::string msg = err0;
    for (int r = 1; r < num_devices && bad == GRAV_B200_OK; r++) if (T->rc[r]) { bad = T->rc[r]; msg = T->err[r]; }
    if (bad != GRAV_B200_OK) {
        grav_b200_ctx *leader = T->ctx[0];
        if (leader) { leader->team = T; team_destroy(leader); grav_b200_ctx_destroy(leader); }
        else {
            { std::lock_guard<std::mutex> lk(T->mu); T->stop = true; }
            T->cv_go.notify_all();
            for (auto &t : T->th) if (t.joinable()) t.join();
            delete T;
        }
        set_error("device team: %s", msg.c_str());
        return bad;
    }
    T->ctx[0]->team = T;
    *out = T->ctx[0];
    return GRAV_B200_OK;
}

// Context for the reference-facing paths: device GRAV_B200_DEVICE (default 0), or a team of GRAV_B200_DEVICES devices
// starting there.
extern "C" int grav_b200_ctx_create_auto(grav_b200_ctx **out)
{
    int base = 0, k = 1;
    if (const char *e = getenv("GRAV_B200_DEVICE")) base = atoi(e);
    if (const char *e = getenv("GRAV_B200_DEVICES")) k = atoi(e);
    if (k <= 1) return grav_b200_ctx_create(out, base, 0, 1, nullptr);
    std::vector<int> dev(k);
    for (int r = 0; r < k; r++) dev[r] = base + r;
    return grav_b200_ctx_create_team(out, k, dev.data());
}

extern "C" int grav_b200_ctx_team_size(const grav_b200_ctx *c) { return (c && c->team) ? c->team->k : 1; }
