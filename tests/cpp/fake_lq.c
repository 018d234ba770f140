/* fake_lq.c -- TEST DOUBLE of the C ABI (include/lq.h) for the host-side plumbing tests: it is LD_PRELOADed
 * in front of liblq.so so that `loop` (driver, worker, observables, evaluators, checkpoint) can be run end to
 * end on a machine without a GPU.  It simulates NOTHING: lq_sweep returns canned numbers that depend on the
 * step counter only, lq_get_state / lq_set_state hand back what they were given.  It lives under tests/ and
 * is never built, installed or loaded by the package (the product has no CPU path: csrc/lq_engine.cu fails
 * with "no CUDA device"). */
#include <stdlib.h>
#include <string.h>
#include "../../include/lq.h"

typedef struct fake {
  int n;
  double beta, site_weight;
  uint32_t step;
  int32_t* spins;
  lq_op* ops;
  int64_t nops;
} fake;

static const char* g_err = "";
const char* lq_last_error(void) { return g_err; }

int lq_create(lq_handle* out, const lq_lattice* lat, const lq_model* model, double beta, const lq_options* opt) {
  (void)opt;
  if (!out || !lat || !model || lat->num_sites <= 0 || !lat->src || !lat->dst) { g_err = "fake: bad arguments"; return -1; }
  fake* f = (fake*)calloc(1, sizeof(fake));
  f->n = lat->num_sites;
  f->beta = beta;
  f->site_weight = model->uniform_site_weight;
  f->spins = (int32_t*)calloc((size_t)f->n, sizeof(int32_t));
  *out = (lq_handle)f;
  return LQ_OK;
}
int lq_destroy(lq_handle h) {
  fake* f = (fake*)h;
  if (f) { free(f->spins); free(f->ops); free(f); }
  return LQ_OK;
}
int lq_set_beta(lq_handle h, double beta) { ((fake*)h)->beta = beta; return LQ_OK; }
int lq_sweep(lq_handle h, lq_collector* c) {
  fake* f = (fake*)h;
  const double x = (double)((f->step * 2654435761u) >> 24) / 256.0;   /* in [0, 1), a function of the step */
  memset(c, 0, sizeof *c);
  c->nop = 10 + (double)(f->step % 5);
  c->nc = 3 + (double)(f->step % 3);
  c->ene = -0.5 * f->n - x;
  c->umag2 = 0.25 * x; c->umag4 = 0.05 * x * x; c->umag = 0.1 * x;
  c->usize2 = 1 + x; c->usize4 = 1 + x * x; c->usize = 0.5 + x;
  c->smag2 = 1 + x; c->smag4 = 1 + x * x; c->smag = 0.5 + x;
  c->ssize2 = 0.25 * x; c->ssize4 = 0.05 * x * x; c->ssize = 0.1 * x;
  c->tlen = f->site_weight > 0 ? 2 * x : 0;
  c->w2 = x;
  ++f->step;
  return LQ_OK;
}
int lq_get_state(lq_handle h, int32_t* spins, lq_op* ops, int64_t* n) {
  fake* f = (fake*)h;
  if (n) *n = f->nops;
  if (spins) memcpy(spins, f->spins, (size_t)f->n * sizeof(int32_t));
  if (ops && f->nops) memcpy(ops, f->ops, (size_t)f->nops * sizeof(lq_op));
  return LQ_OK;
}
int lq_set_state(lq_handle h, const int32_t* spins, const lq_op* ops, int64_t n) {
  fake* f = (fake*)h;
  memcpy(f->spins, spins, (size_t)f->n * sizeof(int32_t));
  free(f->ops);
  f->ops = n ? (lq_op*)malloc((size_t)n * sizeof(lq_op)) : NULL;
  if (n) memcpy(f->ops, ops, (size_t)n * sizeof(lq_op));
  f->nops = n;
  return LQ_OK;
}
uint32_t lq_get_step(lq_handle h) { return ((fake*)h)->step; }
int lq_set_step(lq_handle h, uint32_t step) { ((fake*)h)->step = step; return LQ_OK; }
int lq_get_info(lq_handle h, lq_info* out) {
  (void)h;
  memset(out, 0, sizeof *out);
  out->num_tiles = 1;
  return LQ_OK;
}
/* multi-rank rendezvous of the driver (--nranks P): the id is a token, the "communicator" remembers it */
int lq_comm_unique_id(void* id_out) { memset(id_out, 0x5a, LQ_NCCL_ID_BYTES); return LQ_OK; }
int lq_comm_init(lq_handle h, const void* id, int32_t rank, int32_t nranks) {
  (void)h;
  if (rank < 0 || rank >= nranks || ((const unsigned char*)id)[LQ_NCCL_ID_BYTES - 1] != 0x5a) { g_err = "fake: bad communicator id"; return -1; }
  return LQ_OK;
}
