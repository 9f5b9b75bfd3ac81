// sharded_ipc_test.cpp -- the multi-rank frame-to-map match driven from C++ only (no Python, no torch, no NCCL):
// N processes (one per rank, rank r on GPU r % device_count -- two ranks on ONE GPU work too), communicator set up through
// the C ABI (bshot_comm_create / export / import, handles exchanged over pipes), every rank holds a contiguous shard of
// the same seeded map and checks bshot_match_map_sharded against bshot_match_map of the WHOLE map on a second context:
// records must be identical bit for bit (top-2 with the first-minimum tie-break across shards, rq of the winner).
// usage: sharded_ipc_test [nranks=2] [T=6000] [Q=700]
#include <sys/wait.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../include/bshot_b200.h"

static uint64_t rng_state;
static uint64_t next_u64() {  // splitmix64
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static void fill(std::vector<uint64_t>& d, size_t n, uint64_t seed) {
    rng_state = seed;
    d.resize(n * 6);
    for (size_t i = 0; i < n; ++i) {
        for (int k = 0; k < 6; ++k) d[6 * i + k] = next_u64() & next_u64();  // ~88 of 352 bits set
        d[6 * i + 5] &= 0xFFFFFFFFull;
    }
}

static int run_rank(int rank, int nranks, size_t T, size_t Q, int (*pipes)[8][2]) {
    bshot_ctx *ctx = nullptr, *whole = nullptr;
    int ndev = 1;
    const size_t per = (T + nranks - 1) / nranks, lo = rank * per, hi = std::min(T, lo + per);
    std::vector<uint64_t> map, q;
    fill(map, T, 7);
    fill(q, Q, 8);
    for (size_t i = 0; i < 64 && i < Q; ++i) memcpy(&q[6 * i], &map[6 * ((i * 97) % T)], 48);          // exact copies of map entries
    for (size_t i = 0; i + 1 < T; i += T / 16) memcpy(&map[6 * (T - 1 - i / 2)], &map[6 * i], 48);       // duplicates across shards: lowest index must win
#define TRY(x) do { if ((x) != BSHOT_OK) { std::printf("rank %d: %s failed: %s\n", rank, #x, bshot_last_error()); return 1; } } while (0)
    if (bshot_ctx_create(&ctx, 0, 1024, 1024, T) != BSHOT_OK) { std::printf("rank %d: %s\n", rank, bshot_last_error()); return 1; }
    bshot_ctx_destroy(ctx);
    // device count through a throw-away context is not exposed; ranks share GPU 0 unless BSHOT_TEST_NDEV says otherwise
    if (const char* e = getenv("BSHOT_TEST_NDEV")) ndev = std::max(1, atoi(e));
    TRY(bshot_ctx_create(&ctx, rank % ndev, 1024, 1024, per + 16));
    TRY(bshot_ctx_create(&whole, rank % ndev, 1024, 1024, T));
    TRY(bshot_map_append(ctx, &map[6 * lo], hi - lo));
    TRY(bshot_map_append(whole, map.data(), T));
    TRY(bshot_comm_create(ctx, rank, nranks, Q));
    std::vector<bshot_ipc_handle> handles(nranks);
    TRY(bshot_comm_export(ctx, &handles[rank]));
    for (int p = 0; p < nranks; ++p)
        if (p != rank && write(pipes[rank][p][1], &handles[rank], sizeof(bshot_ipc_handle)) != (ssize_t)sizeof(bshot_ipc_handle)) return 1;
    for (int p = 0; p < nranks; ++p)
        if (p != rank && read(pipes[p][rank][0], &handles[p], sizeof(bshot_ipc_handle)) != (ssize_t)sizeof(bshot_ipc_handle)) return 1;
    TRY(bshot_comm_import(ctx, handles.data()));
    std::vector<bshot_cand> got(Q), want(Q);
    int bad = 0;
    for (int call = 0; call < 3; ++call) {                       // several calls: the flag epochs must keep the ranks in step
        const size_t nq = call == 2 ? Q / 3 : Q;                  // and a smaller query set must not see stale records
        TRY(bshot_match_map_sharded(ctx, q.data(), nq, lo, got.data()));
        TRY(bshot_match_map(whole, q.data(), nq, 0, want.data()));
        for (size_t i = 0; i < nq; ++i)
            if (got[i].k1 != want[i].k1 || got[i].k2 != want[i].k2 || got[i].rq != want[i].rq) {
                if (bad++ < 5) std::printf("rank %d call %d query %zu: got (%llx %llx %u) want (%llx %llx %u)\n", rank, call, i, (unsigned long long)got[i].k1,
                                           (unsigned long long)got[i].k2, got[i].rq, (unsigned long long)want[i].k1, (unsigned long long)want[i].k2, want[i].rq);
            }
    }
    // barrier over the pipes before tearing the regions down (a peer may still be reading ours)
    char tok = 1;
    for (int p = 0; p < nranks; ++p) if (p != rank && write(pipes[rank][p][1], &tok, 1) != 1) return 1;
    for (int p = 0; p < nranks; ++p) if (p != rank && read(pipes[p][rank][0], &tok, 1) != 1) return 1;
    TRY(bshot_comm_destroy(ctx));
    bshot_ctx_destroy(ctx);
    bshot_ctx_destroy(whole);
    if (bad) { std::printf("rank %d: %d records differ\n", rank, bad); return 1; }
    std::printf("rank %d/%d ok: shard [%zu, %zu), %zu queries\n", rank, nranks, lo, hi, Q);
    return 0;
}

int main(int argc, char** argv) {
    const int nranks = argc > 1 ? atoi(argv[1]) : 2;
    const size_t T = argc > 2 ? (size_t)atol(argv[2]) : 6000, Q = argc > 3 ? (size_t)atol(argv[3]) : 700;
    if (nranks < 1 || nranks > 8) return 2;
    static int pipes[8][8][2];
    for (int i = 0; i < nranks; ++i)
        for (int j = 0; j < nranks; ++j)
            if (pipe(pipes[i][j]) != 0) return 2;
    std::vector<pid_t> kids;
    for (int r = 0; r < nranks; ++r) {   // fork BEFORE any CUDA call: every rank initialises CUDA itself
        const pid_t pid = fork();
        if (pid == 0) { const int rc = run_rank(r, nranks, T, Q, pipes); std::fflush(nullptr); _exit(rc); }
        kids.push_back(pid);
    }
    int rc = 0;
    for (pid_t k : kids) {
        int st = 0;
        waitpid(k, &st, 0);
        if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) rc = 1;
    }
    std::printf(rc ? "FAILED\n" : "all ranks ok\n");
    return rc;
}
