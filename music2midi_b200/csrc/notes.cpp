// token stream -> note rows: the integer state machine of MidiTokenizer._decode_tokens and the
// njit helper _tokens_to_note (reference music2midi/tokenizer.py:169-200, 242-267).  CPU code: a few
// thousand branchy integer steps per 3 s segment, no data parallelism worth a launch.
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <vector>

#include "../../include/m2m_b200.h"

namespace m2m {
#ifdef M2M_NOTES_STANDALONE
// host-only build (libm2m_notes.so, g++): the tokenizer works without the CUDA toolchain / a GPU
static thread_local char g_notes_err[256] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_notes_err, sizeof(g_notes_err), fmt, ap);
  va_end(ap);
}
#else
void set_error(const char* fmt, ...);
#endif
}

#ifdef M2M_NOTES_STANDALONE
extern "C" const char* m2m_notes_last_error(void) { return m2m::g_notes_err; }
#endif

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

// Whole token matrix [n_rows, row_len] in one call (Music2MIDI.sample_tokens / generate_many decode thousands of
// segments per batch): row i is decoded independently with start_idx = start_idx0 + i * steps_per_row (the
// "sequential" mode of MidiTokenizer.decode, reference tokenizer.py:75-83; steps_per_row = 0 gives the "batched"
// mode) and the note rows are concatenated; row_note_end[i] = number of note rows after row i.
extern "C" int m2m_tokens_to_notes_batch(const int64_t* tokens, int64_t n_rows, int64_t row_len, int64_t start_idx0,
                                         int64_t steps_per_row, int32_t pitch_offset, int32_t time_offset,
                                         int32_t velocity, int64_t* out_rows4, int64_t cap, int64_t* row_note_end,
                                         int64_t* n_notes) {
  if ((!tokens && n_rows * row_len > 0) || !n_notes || n_rows < 0 || row_len < 0) {
    m2m::set_error("m2m_tokens_to_notes_batch: bad argument");
    return M2M_ERR_INVALID;
  }
  int64_t total = 0;
  for (int64_t i = 0; i < n_rows; ++i) {
    int64_t n = 0;
    int rc = m2m_tokens_to_notes(tokens + i * row_len, row_len, start_idx0 + i * steps_per_row, pitch_offset, time_offset,
                                 velocity, out_rows4 ? out_rows4 + 4 * total : nullptr, cap - total, &n);
    if (rc != M2M_OK) return rc;
    total += n;
    if (row_note_end) row_note_end[i] = total;
  }
  *n_notes = total;
  return M2M_OK;
}
