// Native host runtime for keeping several batches in flight: one worker thread per context.
//
// A step's host work (layout compile ~10 us, table upload, three launches) is about as long as its
// GPU work at batch 64 once several streams overlap, so a single host thread feeding K streams
// caps the throughput. Here every context (= stream) has its own worker thread with a job queue;
// n2nmn_pool_submit only copies the token matrix into a job and returns. There is no reference
// counterpart: the reference's executor is a Python loop around session.run
// (exp_clevr/eval_clevr.py:96-133), one batch at a time.
//
// Uses nothing but the public C ABI of include/n2nmn_b200.h.
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/n2nmn_b200.h"

namespace {

struct Job {
  const float* feat;
  const float* wv;
  std::vector<int32_t> tokens;
  int T, N;
  float* scores;
  uint8_t* validity;
  int host_io;
};

struct Worker {
  n2nmn_ctx* ctx = nullptr;
  void* stream = nullptr;
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::deque<Job> q;
  bool stop = false;
};

}  // namespace

struct n2nmn_pool {
  std::vector<Worker*> workers;
  std::vector<int32_t> vocab;
  std::mutex done_mu;
  std::condition_variable done_cv;
  int64_t pending = 0;        // jobs queued or running
  int err_code = 0;           // first failure since the last wait
  std::string err_msg;
};

namespace {

thread_local std::string g_pool_err;

void run_job(n2nmn_pool* p, Worker* w, Job& j) {
  int rc;
  if (j.host_io) {
    rc = n2nmn_forward_host_async(w->ctx, j.feat, j.wv, j.tokens.data(), j.T, j.N,
                                  p->vocab.data(), (int)p->vocab.size(), j.scores, j.validity,
                                  w->stream);
  } else {
    rc = n2nmn_forward_tokens(w->ctx, j.feat, j.wv, j.tokens.data(), j.T, j.N, p->vocab.data(),
                              (int)p->vocab.size(), j.scores, j.validity, w->stream);
  }
  std::lock_guard<std::mutex> lk(p->done_mu);
  if (rc != 0 && p->err_code == 0) {
    p->err_code = rc;
    p->err_msg = n2nmn_last_error();
  }
  if (--p->pending == 0) p->done_cv.notify_all();
}

void worker_main(n2nmn_pool* p, Worker* w) {
  for (;;) {
    Job j;
    {
      std::unique_lock<std::mutex> lk(w->mu);
      w->cv.wait(lk, [&] { return w->stop || !w->q.empty(); });
      if (w->q.empty()) return;   // stop requested and nothing left
      j = std::move(w->q.front());
      w->q.pop_front();
    }
    run_job(p, w, j);
  }
}

}  // namespace

extern "C" {

const char* n2nmn_pool_last_error(void) { return g_pool_err.c_str(); }

int n2nmn_pool_create(n2nmn_ctx** ctxs, void** streams, int num, const int32_t* vocab_ops,
                      int num_vocab, n2nmn_pool** out) {
  if (!ctxs || !streams || !vocab_ops || !out || num <= 0 || num_vocab <= 0) {
    g_pool_err = "n2nmn_pool_create: bad argument";
    return N2NMN_ERR_ARG;
  }
  n2nmn_pool* p = new n2nmn_pool();
  p->vocab.assign(vocab_ops, vocab_ops + num_vocab);
  for (int i = 0; i < num; ++i) {
    Worker* w = new Worker();
    w->ctx = ctxs[i];
    w->stream = streams[i];
    p->workers.push_back(w);
  }
  for (Worker* w : p->workers) w->th = std::thread(worker_main, p, w);
  *out = p;
  return 0;
}

int n2nmn_pool_destroy(n2nmn_pool* p) {
  if (!p) return 0;
  for (Worker* w : p->workers) {
    {
      std::lock_guard<std::mutex> lk(w->mu);
      w->stop = true;
    }
    w->cv.notify_all();
  }
  for (Worker* w : p->workers) {
    if (w->th.joinable()) w->th.join();
    delete w;
  }
  delete p;
  return 0;
}

int n2nmn_pool_size(const n2nmn_pool* p) { return p ? (int)p->workers.size() : 0; }

int n2nmn_pool_submit(n2nmn_pool* p, int slot, const float* feat, const float* wv,
                      const int32_t* tokens, int T, int N, float* scores, uint8_t* validity_out,
                      int host_io) {
  if (!p || !feat || !wv || !tokens || !scores || slot < 0 || slot >= (int)p->workers.size() ||
      T <= 0 || N <= 0) {
    g_pool_err = "n2nmn_pool_submit: bad argument";
    return N2NMN_ERR_ARG;
  }
  Job j;
  j.feat = feat; j.wv = wv; j.T = T; j.N = N; j.scores = scores; j.validity = validity_out;
  j.host_io = host_io;
  j.tokens.assign(tokens, tokens + (size_t)T * N);
  {
    std::lock_guard<std::mutex> lk(p->done_mu);
    ++p->pending;
  }
  Worker* w = p->workers[slot];
  {
    std::lock_guard<std::mutex> lk(w->mu);
    w->q.push_back(std::move(j));
  }
  w->cv.notify_one();
  return 0;
}

int n2nmn_pool_wait(n2nmn_pool* p) {
  if (!p) return 0;
  std::unique_lock<std::mutex> lk(p->done_mu);
  p->done_cv.wait(lk, [&] { return p->pending == 0; });
  const int rc = p->err_code;
  if (rc != 0) g_pool_err = p->err_msg;
  p->err_code = 0;
  p->err_msg.clear();
  return rc;
}

}  // extern "C"
