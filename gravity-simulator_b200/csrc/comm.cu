// NCCL plumbing: one process per GPU, one communicator per context.
// The only data-path collective is the all-gather of the owned position shards before a
// force evaluation (32 bytes per particle in total); downloads gather v / a shards on demand.
#include <nccl.h>
#include "internal.cuh"

namespace gb {

static int nccl_fail(ncclResult_t r, const char *what)
{
    set_error("NCCL error %d (%s) in %s", (int)r, ncclGetErrorString(r), what);
    return GRAV_B200_ENCCL;
}
#define GB_NCCL(call)                                         \
    do {                                                      \
        ncclResult_t r_ = (call);                             \
        if (r_ != ncclSuccess) return nccl_fail(r_, #call);   \
    } while (0)

int comm_init(grav_b200_ctx *c, const void *uid)
{
    static_assert(sizeof(ncclUniqueId) == 128, "grav_b200.h documents a 128-byte id");
    ncclUniqueId id;
    memcpy(&id, uid, sizeof(id));
    ncclComm_t comm;
    GB_NCCL(ncclCommInitRank(&comm, c->world, id, c->rank));
    c->comm = (ncclComm *)comm;
    return GRAV_B200_OK;
}

void comm_destroy(grav_b200_ctx *c)
{
    if (c->comm) ncclCommDestroy((ncclComm_t)c->comm);
    c->comm = nullptr;
}

// Shards are [r*n/W, (r+1)*n/W): sizes differ by at most one, so use grouped broadcasts
// (NCCL fuses them into one all-gather-like operation) instead of padding.
static int allgather_var(grav_b200_ctx *c, char *base, size_t elem_bytes)
{
    GB_NCCL(ncclGroupStart());
    for (int r = 0; r < c->world; r++) {
        const long long lo = ((long long)r * c->n) / c->world, hi = ((long long)(r + 1) * c->n) / c->world;
        if (hi > lo)
            GB_NCCL(ncclBroadcast(base + lo * elem_bytes, base + lo * elem_bytes, (size_t)(hi - lo) * elem_bytes, ncclChar, r,
                                  (ncclComm_t)c->comm, c->stream));
    }
    GB_NCCL(ncclGroupEnd());
    return GRAV_B200_OK;
}

int comm_allgather_posm(grav_b200_ctx *c)
{
    if (c->world == 1) return GRAV_B200_OK;
    if (c->n % c->world == 0) {
        const size_t cnt = (size_t)(c->n / c->world) * 4;   // doubles per shard
        double *base = c->posm.as<double>();
        GB_NCCL(ncclAllGather(base + (size_t)c->rank * cnt, base, cnt, ncclDouble, (ncclComm_t)c->comm, c->stream));
        return GRAV_B200_OK;
    }
    return allgather_var(c, (char *)c->posm.p, sizeof(double4));
}

int comm_allgather_aos3(grav_b200_ctx *c, double *d_buf)
{
    if (c->world == 1) return GRAV_B200_OK;
    return allgather_var(c, (char *)d_buf, 3 * sizeof(double));
}

int comm_allgather_equal(grav_b200_ctx *c, double *d_base, size_t count_per_rank)
{
    if (c->world == 1 || count_per_rank == 0) return GRAV_B200_OK;
    GB_NCCL(ncclAllGather(d_base + (size_t)c->rank * count_per_rank, d_base, count_per_rank, ncclDouble, (ncclComm_t)c->comm, c->stream));
    return GRAV_B200_OK;
}

int comm_allreduce_sum(grav_b200_ctx *c, double *d_val, int count)
{
    if (c->world == 1) return GRAV_B200_OK;
    GB_NCCL(ncclAllReduce(d_val, d_val, (size_t)count, ncclDouble, ncclSum, (ncclComm_t)c->comm, c->stream));
    return GRAV_B200_OK;
}

}  // namespace gb

extern "C" int grav_b200_nccl_unique_id(void *out128)
{
    if (!out128) { gb::set_error("NULL out pointer"); return GRAV_B200_EINVAL; }
    ncclUniqueId id;
    ncclResult_t r = ncclGetUniqueId(&id);
    if (r != ncclSuccess) return gb::nccl_fail(r, "ncclGetUniqueId");
    memcpy(out128, &id, sizeof(id));
    return GRAV_B200_OK;
}
