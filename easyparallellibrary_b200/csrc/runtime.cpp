// Host runtime helpers (C ABI): pipeline schedule generation + simulation, gradient-bucket
// planning, balanced contiguous partitioning.  These are the planning hot loops that the
// reference runs in Python over TF NodeDefs (graph_editor.py, scheduler.py, partitioner.py);
// here they are native and the Python implementations (parallel/schedule.py,
// communicators/coalescing.py, parallel/partitioner.py) are cross-checked against them in the
// tests.  Opcodes match parallel/schedule.py.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <vector>

namespace {

enum Op : int { F = 0, B = 1, SEND_F = 2, RECV_F = 3, SEND_B = 4, RECV_B = 5, REDUCE = 6, APPLY = 7 };
enum Policy : int { PREFER_FORWARD = 0, PREFER_BACKWARD = 1, PREFER_BACKWARD_OPT = 2 };

struct Instr { int op; int mb; };

int in_flight_cap(int policy, int stage, int S, int M) {
  if (policy == PREFER_FORWARD) return M;
  int extra = (policy == PREFER_BACKWARD_OPT && stage != S - 1) ? 1 : 0;
  return std::max(1, std::min(S - stage + extra, M));
}

std::vector<Instr> stage_program(int policy, int stage, int S, int M, int prefetch) {
  const int cap = in_flight_cap(policy, stage, S, M);
  std::vector<Instr> order;
  for (int f = 0, b = 0; b < M;) {
    if (f < M && f - b < cap) order.push_back({F, f++});
    else order.push_back({B, b++});
  }
  const bool first = stage == 0, last = stage == S - 1;
  struct Slot { std::vector<Instr> pre; Instr ins; std::vector<Instr> post; };
  std::vector<Slot> slots;
  for (const Instr& ins : order) {
    Slot s; s.ins = ins;
    if (ins.op == F) { if (!first) s.pre.push_back({RECV_F, ins.mb}); if (!last) s.post.push_back({SEND_F, ins.mb}); }
    else { if (!last) s.pre.push_back({RECV_B, ins.mb}); if (!first) s.post.push_back({SEND_B, ins.mb}); }
    slots.push_back(s);
  }
  std::vector<Instr> prog;
  std::vector<char> posted(slots.size(), 0);
  std::vector<size_t> f_slot(M, 0);
  for (size_t i = 0; i < slots.size(); ++i) if (slots[i].ins.op == F) f_slot[slots[i].ins.mb] = i;
  for (size_t i = 0; i < slots.size(); ++i) {
    int ahead = 0;
    for (size_t j = i; j < slots.size() && ahead <= std::max(prefetch, 0); ++j) {
      // a gradient receive is never hoisted above the forward of its own micro-batch
      const bool hoistable = j == i || slots[j].ins.op == F || f_slot[slots[j].ins.mb] < i;
      if ((!slots[j].pre.empty() && hoistable) || j == i) {
        if (!posted[j]) { posted[j] = 1; for (const Instr& r : slots[j].pre) prog.push_back(r); }
        if (j > i) ++ahead;
      }
    }
    prog.push_back(slots[i].ins);
    for (const Instr& p : slots[i].post) prog.push_back(p);
  }
  prog.push_back({REDUCE, -1});
  prog.push_back({APPLY, -1});
  return prog;
}

}  // namespace

extern "C" {

// Writes up to `cap` (op, mb) pairs into out[2*i], out[2*i+1]; returns the program length (or -1 if cap is too small).
int epl_schedule_stage(int policy, int stage, int num_stages, int num_micro_batch, int prefetch, int32_t* out, int cap) {
  std::vector<Instr> p = stage_program(policy, stage, num_stages, num_micro_batch, prefetch);
  if ((int)p.size() > cap) return -1;
  for (size_t i = 0; i < p.size(); ++i) { out[2 * i] = p[i].op; out[2 * i + 1] = p[i].mb; }
  return (int)p.size();
}

// Simulates all stages; returns 0 if deadlock-free.  makespan/bubble are outputs; max_in_flight has num_stages ints.
int epl_schedule_simulate(int policy, int S, int M, int prefetch, double t_fwd, double t_bwd, double t_p2p,
                          double* makespan, double* bubble, int32_t* max_in_flight) {
  std::vector<std::vector<Instr>> progs;
  for (int s = 0; s < S; ++s) progs.push_back(stage_program(policy, s, S, M, prefetch));
  std::vector<size_t> pc(S, 0);
  std::vector<double> clock(S, 0.0), busy(S, 0.0);
  std::vector<int> infl(S, 0);
  for (int s = 0; s < S; ++s) max_in_flight[s] = 0;
  std::map<std::pair<int, std::pair<int, int>>, double> sent;
  bool progress = true;
  while (progress) {
    progress = false;
    for (int s = 0; s < S; ++s) {
      while (pc[s] < progs[s].size()) {
        Instr ins = progs[s][pc[s]];
        if (ins.op == RECV_F || ins.op == RECV_B || ins.op == REDUCE || ins.op == APPLY) { ++pc[s]; progress = true; continue; }
        if (ins.op == SEND_F || ins.op == SEND_B) {
          int dst = ins.op == SEND_F ? s + 1 : s - 1;
          sent[{ins.op, {dst, ins.mb}}] = clock[s] + t_p2p;
          ++pc[s]; progress = true; continue;
        }
        bool need = (ins.op == F && s > 0) || (ins.op == B && s < S - 1);
        if (need) {
          auto it = sent.find({ins.op == F ? SEND_F : SEND_B, {s, ins.mb}});
          if (it == sent.end()) break;
          bool posted = false;
          for (size_t k = 0; k < pc[s]; ++k)
            if (progs[s][k].op == (ins.op == F ? RECV_F : RECV_B) && progs[s][k].mb == ins.mb) posted = true;
          if (!posted) return 2;
          clock[s] = std::max(clock[s], it->second);
        }
        double dur = ins.op == F ? t_fwd : t_bwd;
        clock[s] += dur; busy[s] += dur;
        infl[s] += ins.op == F ? 1 : -1;
        max_in_flight[s] = std::max(max_in_flight[s], infl[s]);
        ++pc[s]; progress = true;
      }
    }
  }
  for (int s = 0; s < S; ++s) if (pc[s] < progs[s].size()) return 1;
  double mk = 0, tot = 0;
  for (int s = 0; s < S; ++s) { mk = std::max(mk, clock[s]); tot += busy[s]; }
  *makespan = mk;
  *bubble = mk > 0 ? 1.0 - tot / (mk * S) : 0.0;
  return 0;
}

// Bucket planner: same policy as communicators/coalescing.plan_buckets.  dtype_ids are small ints.
// Writes bucket index per tensor into out_bucket[n]; returns number of buckets.
int epl_plan_buckets(const int64_t* nbytes, const int32_t* dtype_ids, int n, int max_splits, int32_t* out_bucket) {
  if (n == 0) return 0;
  if (n == 1) { out_bucket[0] = 0; return 1; }
  std::vector<int> order_dtypes;
  std::map<int, std::vector<int>> groups;
  for (int i = 0; i < n; ++i) {
    if (!groups.count(dtype_ids[i])) order_dtypes.push_back(dtype_ids[i]);
    groups[dtype_ids[i]].push_back(i);
  }
  int nb = 0;
  if ((int)order_dtypes.size() >= max_splits) {
    for (int dt : order_dtypes) { for (int i : groups[dt]) out_bucket[i] = nb; ++nb; }
    return nb;
  }
  std::vector<int> budget;
  long total_span = 0;
  for (int dt : order_dtypes) total_span += std::max<long>((long)groups[dt].size() - 1, 0);
  if (total_span == 0) total_span = 1;
  int sum = 0;
  for (int dt : order_dtypes) {
    long span = std::max<long>((long)groups[dt].size() - 1, 0);
    int b = std::max((int)(max_splits * span / total_span), 1);
    budget.push_back(b); sum += b;
  }
  budget[0] = std::max(budget[0] + (max_splits - sum), 1);
  for (size_t g = 0; g < order_dtypes.size(); ++g) {
    const std::vector<int>& idx = groups[order_dtypes[g]];
    double nz = 0; int nzc = 0;
    for (int i : idx) if (nbytes[i]) { nz += (double)nbytes[i]; ++nzc; }
    double mean = nzc ? nz / nzc : 1.0;
    double tot = 0;
    std::vector<double> sz;
    for (int i : idx) { double s = nbytes[i] ? (double)nbytes[i] : mean; sz.push_back(s); tot += s; }
    double limit = budget[g] == 1 ? tot : tot / (budget[g] - 1);
    double acc = 0; bool open = false;
    for (size_t k = 0; k < idx.size(); ++k) {
      if (open && acc + sz[k] > limit) { ++nb; acc = 0; open = false; }
      out_bucket[idx[k]] = nb; acc += sz[k]; open = true;
    }
    if (open) ++nb;
  }
  return nb;
}

// Min-max contiguous partition into exactly `parts` groups: out_start[g] = first index of group g (out_start[parts] = n).
int epl_partition_stages(const double* weights, int n, int parts, int32_t* out_start) {
  if (parts <= 0) return -1;
  if (n <= parts) {
    for (int g = 0; g <= parts; ++g) out_start[g] = std::min(g, n);
    return 0;
  }
  auto fits = [&](double bound) {
    int used = 1; double cur = 0;
    for (int i = 0; i < n; ++i) {
      if (weights[i] > bound) return false;
      if (cur + weights[i] > bound) { ++used; cur = weights[i]; if (used > parts) return false; }
      else cur += weights[i];
    }
    return true;
  };
  double lo = 0, hi = 0;
  for (int i = 0; i < n; ++i) { lo = std::max(lo, weights[i]); hi += weights[i]; }
  for (int it = 0; it < 64; ++it) { double mid = (lo + hi) / 2; if (fits(mid)) hi = mid; else lo = mid; }
  std::vector<int> starts{0};
  double cur = 0;
  for (int i = 0; i < n; ++i) {
    if (i > starts.back() && cur + weights[i] > hi) { starts.push_back(i); cur = 0; }
    cur += weights[i];
  }
  starts.push_back(n);
  while ((int)starts.size() - 1 < parts) {          // split the heaviest multi-item group
    int best = -1; double bw = -1;
    for (size_t g = 0; g + 1 < starts.size(); ++g) {
      if (starts[g + 1] - starts[g] > 1) {
        double w = 0; for (int i = starts[g]; i < starts[g + 1]; ++i) w += weights[i];
        if (w > bw) { bw = w; best = (int)g; }
      }
    }
    int a = starts[best], b = starts[best + 1], cut = a + 1;
    double acc = 0;
    for (int j = a; j < b - 1; ++j) { acc += weights[j]; cut = j + 1; if (acc >= bw / 2) break; }
    starts.insert(starts.begin() + best + 1, cut);
  }
  for (int g = 0; g <= parts; ++g) out_start[g] = starts[g];
  return 0;
}

}  // extern "C"
