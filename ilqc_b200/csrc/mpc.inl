// ilqc_mpc: B concurrent closed-loop MPC episodes, the window loop of run_mpc on the device side of the C ABI.
//
// Reference semantics restated (file:line in /root/reference):
//   run_mpc   model_predictive_control/mpc_pipeline.py:87-115   horizon <- min(max(horizon + sliding, window), max_horizon)
//   mpc_step  model_predictive_control/mpc_pipeline.py:18-55    warm start, re-anchoring, splice
// With prev_cmd of length `curr` (Python slicing semantics, negative indices included):
//   tail       = prev_cmd[curr - overlap:]            -> rows [tail_start, curr)
//   additional = prev_cmd[-1] repeated (or zeros)      -> full_horizon - curr rows
//   init_cmd   = tail ++ additional
//   init_state = traj(prev_cmd)[curr - overlap]        -> trajectory index a_pos (negative index wraps on curr + 1 states)
//   init_time_iter = curr - overlap                    (may be negative, like in the reference)
//   cmd_opt    = prev_cmd[:-overlap] ++ cmd            -> rows [0, keep) stay, the window's result follows
// In the engine's [t][nu][B] layout every time slice is one contiguous block, so the shift and the splice are
// device-to-device copies, the padding is one small kernel, and the re-anchoring state comes from the roll-out
// kernel advanced over the newly frozen controls only (the reference re-simulates the whole past every window).
#pragma once
#include "solver.inl"

namespace ilqc {
namespace {
// rows [row0, row0 + n_rows) of dst <- the row `src_row` of src (keep_last) or zeros; one row = row_len doubles
ILQC_GLOBAL mpc_pad_kernel(int n, int row_len, const double* src_row, double* dst, int keep_last) {
  const int i = ILQC_TID;
  if (i >= n) return;
  dst[i] = keep_last ? src_row[i % row_len] : 0.0;
}
}  // namespace

struct MpcPlan {
  int first_h;  // horizon of the first window (m->horizon)
  int max_window_h;  // longest solve horizon over all windows
};
inline int mpc_next_horizon(int horizon, int window_size, int overlap, int full_horizon) {
  const int sliding = window_size - overlap;
  int h = horizon + sliding;
  if (h < window_size) h = window_size;
  return h < full_horizon ? h : full_horizon;
}
// geometry of the window that follows a command history of length `curr` and targets `horizon`
struct MpcWindow {
  int tail_start, n_tail, extra, Hw, keep, a, a_pos;
};
inline bool mpc_window(int curr, int horizon, int overlap, bool full, MpcWindow* w) {
  w->extra = horizon - curr;
  if (w->extra < 0) return false;
  if (full) {
    w->tail_start = 0; w->n_tail = curr; w->Hw = horizon; w->keep = 0; w->a = 0; w->a_pos = 0;
    return true;
  }
  const int a = curr - overlap;
  w->a = a;
  w->tail_start = a >= 0 ? a : (curr + a > 0 ? curr + a : 0);   // prev_cmd[a:]
  w->n_tail = curr - w->tail_start;
  w->Hw = w->n_tail + w->extra;
  w->keep = a > 0 ? a : 0;                                       // len(prev_cmd[:-overlap])
  if (overlap == 0) w->keep = 0;                                 // prev_cmd[:-0] is empty in Python
  w->a_pos = a >= 0 ? a : curr + 1 + a;                          // traj[a] on curr + 1 states
  return w->a_pos >= 0 && w->a_pos <= curr;                      // (the reference raises IndexError otherwise)
}

}  // namespace ilqc
