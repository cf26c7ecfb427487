// token stream -> note rows: the integer state machine of MidiTokenizer._decode_tokens and the
// njit helper _tokens_to_note (reference music2midi/tokenizer.py:169-200, 242-267).  CPU code: a few
// thousand branchy integer steps per 3 s segment, no data parallelism worth a launch.
#include <stdint.h>

#include <vector>

#include "../../include/m2m_b200.h"

namespace m2m {
void set_error(const char* fmt, ...);
}

extern "C" int m2m_tokens_to_notes(const int64_t* tokens, int64_t n_tokens, int64_t start_idx, int32_t pitch_offset,
                                   int32_t time_offset, int32_t velocity, int64_t* out_rows4, int64_t cap,
                                   int64_t* n_notes) {
  if ((!tokens && n_tokens > 0) || !n_notes || (!out_rows4 && cap > 0) || n_tokens < 0) {
    m2m::set_error("m2m_tokens_to_notes: bad argument");
    return M2M_ERR_INVALID;
  }
  enum { PAD = 0, BOS = 1, EOS = 2, ONSET = 3, OFFSET = 4 };
  int64_t n = 0;
  int64_t cur_time = -1, cur_on = -1, cur_note = -1;
  // open notes per pitch (indices of rows whose offset is still -1), so that an OFFSET closes its
  // notes without rescanning every row (the reference rescans all rows: O(n^2)).
  std::vector<std::vector<int64_t>> open;
  for (int64_t i = 0; i < n_tokens; ++i) {
    const int64_t tok = tokens[i];
    if (tok == EOS) break;
    if (tok == BOS || tok == PAD) continue;
    if (tok == ONSET) cur_on = 1;
    if (tok == OFFSET) cur_on = 0;
    if (tok >= time_offset) {
      cur_time = start_idx + tok - time_offset;
      cur_on = -1;
      cur_note = -1;
    } else if (tok >= pitch_offset) {
      cur_note = tok - pitch_offset;
    }
    if (cur_time == -1 || cur_on == -1 || cur_note == -1) continue;
    if (cur_on == 1 && velocity != 0) {
      if (n >= cap) {
        m2m::set_error("m2m_tokens_to_notes: output capacity %lld too small", (long long)cap);
        return M2M_ERR_INVALID;
      }
      int64_t* r = out_rows4 + 4 * n;
      r[0] = cur_time;
      r[1] = -1;
      r[2] = cur_note;
      r[3] = velocity;
      if ((size_t)cur_note >= open.size()) open.resize((size_t)cur_note + 1);
      open[(size_t)cur_note].push_back(n);
      ++n;
    } else {
      // note off: every still-open row of this pitch whose onset is STRICTLY earlier gets closed
      if ((size_t)cur_note < open.size()) {
        std::vector<int64_t>& o = open[(size_t)cur_note];
        size_t keep = 0;
        for (size_t k = 0; k < o.size(); ++k) {
          int64_t* r = out_rows4 + 4 * o[k];
          if (r[0] < cur_time)
            r[1] = cur_time;
          else
            o[keep++] = o[k];
        }
        o.resize(keep);
      }
    }
    cur_note = -1;
  }
  *n_notes = n;
  return M2M_OK;
}
