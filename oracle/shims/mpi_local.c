/* Local implementation of the MPI subset declared in mpi.h.
 * Test infrastructure only (used to build/run the PanguLU reference as an
 * oracle and CPU baseline); never linked into the product library.
 *
 * Rank count comes from the environment:
 *   PANGULU_SHIM_NP   (default 1), PANGULU_SHIM_RANK (default 0),
 *   PANGULU_SHIM_SESSION (name of the shared control segment).
 * With NP==1 every collective is a local copy and point-to-point calls trap
 * (the reference never issues them with one rank: every send loop runs over
 * remote ranks and rank_remain_recv_block_count==0).
 * With NP>1 the ranks are separate processes started by the test/bench
 * launcher; messages travel through /dev/shm: one control segment holding a
 * process-shared mutex/condvar, a sense-reversing barrier and one descriptor
 * ring per destination rank, plus one payload file per in-flight message.
 */
#define _GNU_SOURCE
#include "mpi.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <fcntl.h>
#include <errno.h>
#include <pthread.h>
#include <sys/mman.h>
#include <sys/stat.h>

#define SHIM_MAX_RANKS 16
#define SHIM_RING 8192

typedef struct {
    int src, tag;
    long long bytes;
    unsigned long long seq;   /* payload file id */
    int valid;
} shim_desc_t;

typedef struct {
    pthread_mutex_t mu;
    pthread_cond_t cv;
    int np;
    int bar_count;
    int bar_sense;
    unsigned long long next_seq;
    /* per destination: ordered list of pending descriptors (compact array) */
    int nmsg[SHIM_MAX_RANKS];
    shim_desc_t ring[SHIM_MAX_RANKS][SHIM_RING];
} shim_ctrl_t;

static int g_np = 1, g_rank = 0, g_inited = 0;
static shim_ctrl_t *g_ctrl = NULL;
static char g_session[128] = "pangulu_shim";
static pthread_mutex_t g_init_mu = PTHREAD_MUTEX_INITIALIZER;

static void shim_die(const char *what)
{
    fprintf(stderr, "[pangulu mpi shim] %s (rank %d/%d)\n", what, g_rank, g_np);
    abort();
}

static void shim_init(void)
{
    pthread_mutex_lock(&g_init_mu);
    if (g_inited) { pthread_mutex_unlock(&g_init_mu); return; }
    const char *s;
    if ((s = getenv("PANGULU_SHIM_NP"))) g_np = atoi(s);
    if ((s = getenv("PANGULU_SHIM_RANK"))) g_rank = atoi(s);
    if ((s = getenv("PANGULU_SHIM_SESSION"))) snprintf(g_session, sizeof g_session, "%s", s);
    if (g_np < 1 || g_np > SHIM_MAX_RANKS || g_rank < 0 || g_rank >= g_np) shim_die("bad PANGULU_SHIM_NP/RANK");
    if (g_np > 1) {
        char name[200];
        snprintf(name, sizeof name, "/%s_ctrl", g_session);
        int fd = shm_open(name, O_RDWR, 0600);
        if (fd < 0) shim_die("control segment missing: the launcher must create it (see oracle/shim_launch.py)");
        g_ctrl = (shim_ctrl_t *)mmap(NULL, sizeof(shim_ctrl_t), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (g_ctrl == MAP_FAILED) shim_die("mmap failed");
    }
    g_inited = 1;
    pthread_mutex_unlock(&g_init_mu);
}

/* Called once by the launcher process before forking ranks. */
int pangulu_shim_create_session(const char *session, int np)
{
    char name[200];
    snprintf(name, sizeof name, "/%s_ctrl", session);
    shm_unlink(name);
    int fd = shm_open(name, O_CREAT | O_RDWR, 0600);
    if (fd < 0) return -1;
    if (ftruncate(fd, sizeof(shim_ctrl_t)) != 0) { close(fd); return -2; }
    shim_ctrl_t *c = (shim_ctrl_t *)mmap(NULL, sizeof(shim_ctrl_t), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (c == MAP_FAILED) return -3;
    memset(c, 0, sizeof *c);
    pthread_mutexattr_t ma; pthread_mutexattr_init(&ma); pthread_mutexattr_setpshared(&ma, PTHREAD_PROCESS_SHARED);
    pthread_mutex_init(&c->mu, &ma);
    pthread_condattr_t ca; pthread_condattr_init(&ca); pthread_condattr_setpshared(&ca, PTHREAD_PROCESS_SHARED);
    pthread_cond_init(&c->cv, &ca);
    c->np = np;
    munmap(c, sizeof *c);
    return 0;
}

int pangulu_shim_destroy_session(const char *session)
{
    char name[200];
    snprintf(name, sizeof name, "/%s_ctrl", session);
    return shm_unlink(name);
}

static void payload_name(char *out, size_t cap, unsigned long long seq)
{
    snprintf(out, cap, "/%s_m%llu", g_session, seq);
}

static void shim_send(const void *buf, long long bytes, int dest, int tag)
{
    if (g_np == 1) shim_die("point-to-point send with a single rank");
    if (dest < 0 || dest >= g_np) shim_die("send: bad destination");
    pthread_mutex_lock(&g_ctrl->mu);
    unsigned long long seq = g_ctrl->next_seq++;
    pthread_mutex_unlock(&g_ctrl->mu);
    if (bytes > 0) {
        char name[256];
        payload_name(name, sizeof name, seq);
        int fd = shm_open(name, O_CREAT | O_RDWR | O_EXCL, 0600);
        if (fd < 0) shim_die("send: shm_open payload failed");
        if (ftruncate(fd, bytes) != 0) shim_die("send: ftruncate failed");
        void *p = mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (p == MAP_FAILED) shim_die("send: mmap payload failed");
        memcpy(p, buf, bytes);
        munmap(p, bytes);
    }
    pthread_mutex_lock(&g_ctrl->mu);
    while (g_ctrl->nmsg[dest] >= SHIM_RING) pthread_cond_wait(&g_ctrl->cv, &g_ctrl->mu);
    shim_desc_t *d = &g_ctrl->ring[dest][g_ctrl->nmsg[dest]++];
    d->src = g_rank; d->tag = tag; d->bytes = bytes; d->seq = seq; d->valid = 1;
    pthread_cond_broadcast(&g_ctrl->cv);
    pthread_mutex_unlock(&g_ctrl->mu);
}

/* find first pending descriptor for me matching (src,tag); caller holds the mutex */
static int shim_find(int src, int tag)
{
    int n = g_ctrl->nmsg[g_rank];
    for (int i = 0; i < n; i++) {
        shim_desc_t *d = &g_ctrl->ring[g_rank][i];
        if ((src == MPI_ANY_SOURCE || d->src == src) && (tag == MPI_ANY_TAG || d->tag == tag)) return i;
    }
    return -1;
}

static void shim_recv(void *buf, long long cap, int src, int tag, MPI_Status *st)
{
    if (g_np == 1) shim_die("point-to-point recv with a single rank");
    pthread_mutex_lock(&g_ctrl->mu);
    int i;
    while ((i = shim_find(src, tag)) < 0) pthread_cond_wait(&g_ctrl->cv, &g_ctrl->mu);
    shim_desc_t d = g_ctrl->ring[g_rank][i];
    int n = g_ctrl->nmsg[g_rank];
    memmove(&g_ctrl->ring[g_rank][i], &g_ctrl->ring[g_rank][i + 1], sizeof(shim_desc_t) * (size_t)(n - i - 1));
    g_ctrl->nmsg[g_rank] = n - 1;
    pthread_cond_broadcast(&g_ctrl->cv);
    pthread_mutex_unlock(&g_ctrl->mu);
    if (d.bytes > cap) shim_die("recv: message longer than buffer");
    if (d.bytes > 0) {
        char name[256];
        payload_name(name, sizeof name, d.seq);
        int fd = shm_open(name, O_RDWR, 0600);
        if (fd < 0) shim_die("recv: payload missing");
        void *p = mmap(NULL, d.bytes, PROT_READ, MAP_SHARED, fd, 0);
        close(fd);
        if (p == MAP_FAILED) shim_die("recv: mmap payload failed");
        memcpy(buf, p, d.bytes);
        munmap(p, d.bytes);
        shm_unlink(name);
    }
    if (st) { st->MPI_SOURCE = d.src; st->MPI_TAG = d.tag; st->MPI_ERROR = 0; st->_count = d.bytes; }
}

/* internal tags for collectives live far away from the reference's user tags */
#define TAG_BCAST  0x7F000001
#define TAG_REDUCE 0x7F000002

int MPI_Init_thread(int *argc, char ***argv, int required, int *provided)
{
    (void)argc; (void)argv; (void)required;
    shim_init();
    if (provided) *provided = MPI_THREAD_MULTIPLE;
    return MPI_SUCCESS;
}
int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm c, int *rank) { (void)c; shim_init(); *rank = g_rank; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm c, int *size) { (void)c; shim_init(); *size = g_np; return MPI_SUCCESS; }

int MPI_Barrier(MPI_Comm c)
{
    (void)c; shim_init();
    if (g_np == 1) return MPI_SUCCESS;
    pthread_mutex_lock(&g_ctrl->mu);
    int sense = g_ctrl->bar_sense;
    if (++g_ctrl->bar_count == g_np) {
        g_ctrl->bar_count = 0;
        g_ctrl->bar_sense = !sense;
        pthread_cond_broadcast(&g_ctrl->cv);
    } else {
        while (g_ctrl->bar_sense == sense) pthread_cond_wait(&g_ctrl->cv, &g_ctrl->mu);
    }
    pthread_mutex_unlock(&g_ctrl->mu);
    return MPI_SUCCESS;
}

int MPI_Bcast(void *buf, int count, MPI_Datatype dt, int root, MPI_Comm c)
{
    (void)c; shim_init();
    if (g_np == 1) return MPI_SUCCESS;
    long long bytes = (long long)count * dt;
    if (g_rank == root) {
        for (int r = 0; r < g_np; r++) if (r != root) shim_send(buf, bytes, r, TAG_BCAST);
    } else {
        shim_recv(buf, bytes, root, TAG_BCAST, NULL);
    }
    return MPI_SUCCESS;
}

int MPI_Send(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm c)
{ (void)c; shim_init(); shim_send(buf, (long long)count * dt, dest, tag); return MPI_SUCCESS; }

int MPI_Recv(void *buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm c, MPI_Status *st)
{ (void)c; shim_init(); shim_recv(buf, (long long)count * dt, src, tag, st); return MPI_SUCCESS; }

/* eager: the payload is copied out before Isend returns, so the request is born complete */
int MPI_Isend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm c, MPI_Request *req)
{ (void)c; shim_init(); shim_send(buf, (long long)count * dt, dest, tag); if (req) *req = MPI_REQUEST_NULL; return MPI_SUCCESS; }

int MPI_Irecv(void *buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm c, MPI_Request *req)
{
    (void)c; shim_init();
    MPI_Request r = (MPI_Request)calloc(1, sizeof(*r));
    r->kind = 1; r->buf = buf; r->bytes = (long long)count * dt; r->peer = src; r->tag = tag; r->done = 0;
    *req = r;
    return MPI_SUCCESS;
}

int MPI_Wait(MPI_Request *req, MPI_Status *st)
{
    if (!req || *req == MPI_REQUEST_NULL) return MPI_SUCCESS;
    MPI_Request r = *req;
    if (r->kind == 1 && !r->done) shim_recv(r->buf, r->bytes, r->peer, r->tag, st);
    free(r);
    *req = MPI_REQUEST_NULL;
    return MPI_SUCCESS;
}

int MPI_Waitall(int n, MPI_Request *reqs, MPI_Status *sts)
{
    for (int i = 0; i < n; i++) MPI_Wait(&reqs[i], sts ? &sts[i] : NULL);
    return MPI_SUCCESS;
}

int MPI_Iprobe(int src, int tag, MPI_Comm c, int *flag, MPI_Status *st)
{
    (void)c; shim_init();
    *flag = 0;
    if (g_np == 1) return MPI_SUCCESS;
    pthread_mutex_lock(&g_ctrl->mu);
    int i = shim_find(src, tag);
    if (i >= 0) {
        shim_desc_t *d = &g_ctrl->ring[g_rank][i];
        *flag = 1;
        if (st) { st->MPI_SOURCE = d->src; st->MPI_TAG = d->tag; st->MPI_ERROR = 0; st->_count = d->bytes; }
    }
    pthread_mutex_unlock(&g_ctrl->mu);
    return MPI_SUCCESS;
}

int MPI_Probe(int src, int tag, MPI_Comm c, MPI_Status *st)
{
    (void)c; shim_init();
    if (g_np == 1) shim_die("MPI_Probe with a single rank");
    pthread_mutex_lock(&g_ctrl->mu);
    int i;
    while ((i = shim_find(src, tag)) < 0) pthread_cond_wait(&g_ctrl->cv, &g_ctrl->mu);
    shim_desc_t *d = &g_ctrl->ring[g_rank][i];
    if (st) { st->MPI_SOURCE = d->src; st->MPI_TAG = d->tag; st->MPI_ERROR = 0; st->_count = d->bytes; }
    pthread_mutex_unlock(&g_ctrl->mu);
    return MPI_SUCCESS;
}

int MPI_Get_count(const MPI_Status *st, MPI_Datatype dt, int *count)
{ *count = (int)(st->_count / dt); return MPI_SUCCESS; }

/* The reference only reduces with MPI_SUM over 8-byte integers (flop counts,
 * memory totals: src/pangulu.c:243, src/pangulu_utils.c:434,445). */
static void shim_sum(void *acc, const void *x, int count, MPI_Datatype dt)
{
    if (dt != 8) shim_die("reduce: only 8-byte integer SUM is implemented");
    long long *a = (long long *)acc; const long long *b = (const long long *)x;
    for (int i = 0; i < count; i++) a[i] += b[i];
}

int MPI_Reduce(const void *s, void *r, int count, MPI_Datatype dt, MPI_Op op, int root, MPI_Comm c)
{
    (void)c; (void)op; shim_init();
    long long bytes = (long long)count * dt;
    if (g_np == 1) { if (r != s) memcpy(r, s, bytes); return MPI_SUCCESS; }
    if (g_rank == root) {
        if (r != s) memcpy(r, s, bytes);
        void *tmp = malloc(bytes);
        for (int k = 0; k < g_np; k++) if (k != root) { shim_recv(tmp, bytes, k, TAG_REDUCE, NULL); shim_sum(r, tmp, count, dt); }
        free(tmp);
    } else {
        shim_send(s, bytes, root, TAG_REDUCE);
    }
    return MPI_SUCCESS;
}

int MPI_Allreduce(const void *s, void *r, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm c)
{
    MPI_Reduce(s, r, count, dt, op, 0, c);
    MPI_Bcast(r, count, dt, 0, c);
    return MPI_SUCCESS;
}
