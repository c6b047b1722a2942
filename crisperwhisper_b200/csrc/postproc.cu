// postproc.cu — host-only (no device code): token ids + token timestamps -> word chunks.
//
// Native counterpart of crisperwhisper_b200/decode_asr.py for the hot configuration of the reference call
// (`return_timestamps="word"`, byte-level BPE vocabulary), i.e. of tokenizer._decode_asr
// (HF/models/whisper/tokenization_whisper.py:901-1150) with _find_longest_common_sequence (:1153-1270),
// _collate_word_timestamps / _combine_tokens_into_words (:1273-1318), _split_tokens_on_unicode (:1321-1350),
// _split_tokens_on_spaces (:1353-1376) and _merge_punctuations (:1379-1405).  Same algorithm as the Python module (which
// tests/test_decode_asr.py pins to the HF functions); tests/test_postproc_native_cpu.py pins this file to the Python module
// on the same randomised streams.  Whatever the Python path would answer with an exception, and every input this file
// does not model (tokens without a byte spelling), is reported as CW_POST_PUNT: the caller then runs the Python path.
//
// Python semantics reproduced on purpose: str.strip() over Unicode white space, substring `in` tests (the empty string
// is in every string), round(x, 2) as correctly rounded decimal -> double, UTF-8 decoding with errors="replace"
// (one U+FFFD per maximal invalid subpart), character (not byte) offsets in the unicode split.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/crisper.h"

namespace {

typedef std::u32string ustr;
const char32_t kRepl = 0xFFFD;

struct Stamp { double a, b; };

double py_round2(double x) {  // Python round(x, 2): shortest-correct decimal rounding to two places, then back
  if (!isfinite(x)) return x;
  char buf[64];
  snprintf(buf, sizeof buf, "%.2f", x);
  return strtod(buf, nullptr);
}

void utf8_decode_replace(const uint8_t* p, size_t n, ustr& out) {
  size_t i = 0;
  while (i < n) {
    const uint8_t b = p[i];
    int need;
    uint8_t lo = 0x80, hi = 0xBF;
    char32_t cp;
    if (b < 0x80) { out.push_back(b); ++i; continue; }
    else if (b >= 0xC2 && b <= 0xDF) { need = 1; cp = b & 0x1F; }
    else if (b == 0xE0) { need = 2; lo = 0xA0; cp = b & 0x0F; }
    else if ((b >= 0xE1 && b <= 0xEC) || b == 0xEE || b == 0xEF) { need = 2; cp = b & 0x0F; }
    else if (b == 0xED) { need = 2; hi = 0x9F; cp = b & 0x0F; }
    else if (b == 0xF0) { need = 3; lo = 0x90; cp = b & 0x07; }
    else if (b >= 0xF1 && b <= 0xF3) { need = 3; cp = b & 0x07; }
    else if (b == 0xF4) { need = 3; hi = 0x8F; cp = b & 0x07; }
    else { out.push_back(kRepl); ++i; continue; }
    size_t j = i + 1;
    bool ok = true;
    for (int k = 0; k < need; ++k) {
      if (j >= n) { ok = false; break; }
      const uint8_t c = p[j];
      if (c < lo || c > hi) { ok = false; break; }
      cp = (cp << 6) | (c & 0x3F);
      ++j;
      lo = 0x80; hi = 0xBF;
    }
    out.push_back(ok ? cp : kRepl);
    i = j;  // an offending byte is not consumed: it starts the next sequence
  }
}

void utf8_encode(const ustr& s, std::string& out) {
  for (char32_t c : s) {
    if (c < 0x80) out.push_back((char)c);
    else if (c < 0x800) { out.push_back((char)(0xC0 | (c >> 6))); out.push_back((char)(0x80 | (c & 0x3F))); }
    else if (c < 0x10000) {
      out.push_back((char)(0xE0 | (c >> 12))); out.push_back((char)(0x80 | ((c >> 6) & 0x3F))); out.push_back((char)(0x80 | (c & 0x3F)));
    } else {
      out.push_back((char)(0xF0 | (c >> 18))); out.push_back((char)(0x80 | ((c >> 12) & 0x3F)));
      out.push_back((char)(0x80 | ((c >> 6) & 0x3F))); out.push_back((char)(0x80 | (c & 0x3F)));
    }
  }
}

bool py_isspace(char32_t c) {
  return (c >= 0x09 && c <= 0x0D) || (c >= 0x1C && c <= 0x20) || c == 0x85 || c == 0xA0 || c == 0x1680 ||
         (c >= 0x2000 && c <= 0x200A) || c == 0x2028 || c == 0x2029 || c == 0x202F || c == 0x205F || c == 0x3000;
}
ustr py_strip(const ustr& s) {
  size_t a = 0, b = s.size();
  while (a < b && py_isspace(s[a])) ++a;
  while (b > a && py_isspace(s[b - 1])) --b;
  return s.substr(a, b - a);
}
bool py_in(const ustr& needle, const ustr& hay) { return hay.find(needle) != ustr::npos; }  // "" in x is True

const ustr kAsciiPunct = U"!\"#$%&'()*+,-./:;<=>?@[\\]^_`{|}~";
const ustr kPrepended = U"\"'\u201C\u00A1\u00BF([{-";
const ustr kAppended = U"\"'.\u3002,\uFF0C!\uFF01?\uFF1F:\uFF1A\u201D)]}\u3001";

struct Vocab {
  const uint8_t* bytes; const int64_t* off; const uint8_t* has_bytes; int32_t eos_id;
  const uint8_t* is_special; const int32_t* lang_of; int32_t n_ids;
  const uint8_t* lang_unspaced; int32_t n_lang;
};

// Stitch consecutive token runs whose ends overlap in audio (tokenization_whisper.py:1153-1270), with the token
// timestamps (the word-mode variant): see decode_asr.merge_overlaps.
void merge_overlaps(const std::vector<std::vector<int32_t>>& seqs, const std::vector<std::vector<Stamp>>& sts,
                    std::vector<int32_t>& total, std::vector<Stamp>& total_st) {
  total.clear(); total_st.clear();
  const bool use_stamps = !sts.empty();
  std::vector<int32_t> left = seqs[0];
  std::vector<Stamp> left_st;
  if (use_stamps) left_st = sts[0];
  std::vector<int64_t> counts;
  for (size_t si = 1; si < seqs.size(); ++si) {
    const std::vector<int32_t>& right = seqs[si];
    const std::vector<Stamp>* right_st = use_stamps ? &sts[si] : nullptr;
    const int64_t L = (int64_t)left.size(), R = (int64_t)right.size();
    int64_t p0 = L, p1 = L, p2 = 0, p3 = 0;
    if (L && R) {
      counts.assign((size_t)(L + R), 0);
      for (int64_t l = 0; l < L; ++l)
        for (int64_t r = 0; r < R; ++r) {
          if (left[(size_t)l] != right[(size_t)r]) continue;
          if (use_stamps) {
            const Stamp& x = left_st[(size_t)l];
            const Stamp& y = (*right_st)[(size_t)r];
            if (!((x.a < y.a) || (x.a == y.a && x.b <= y.b))) continue;
          }
          counts[(size_t)(r - l + L)] += 1;   // shift i pairs left[l] with right[r] where r - l == i - L
        }
      double best = -1.0;
      int64_t bi = -1;
      for (int64_t i = 1; i < L + R; ++i) {
        if (counts[(size_t)i] <= 1) continue;
        const double score = (double)counts[(size_t)i] / (double)i + (double)i / 10000.0;
        if (score > best) { best = score; bi = i; }
      }
      if (bi >= 0) {
        const int64_t i = bi;
        p0 = (L - i > 0) ? L - i : 0;
        p1 = (L + R - i < L) ? L + R - i : L;
        p2 = (i - L > 0) ? i - L : 0;
        p3 = (i < R) ? i : R;
      }
    }
    const int64_t l_mid = (p1 + p0) / 2, r_mid = (p3 + p2) / 2;
    total.insert(total.end(), left.begin(), left.begin() + l_mid);
    if (use_stamps) total_st.insert(total_st.end(), left_st.begin(), left_st.begin() + l_mid);
    left.assign(right.begin() + r_mid, right.end());
    if (use_stamps) left_st.assign(right_st->begin() + r_mid, right_st->end());
  }
  total.insert(total.end(), left.begin(), left.end());
  if (use_stamps) total_st.insert(total_st.end(), left_st.begin(), left_st.end());
}

struct Word { ustr text; size_t first, last; };   // token index range [first, last] inside the chunk's merged run

// (words, word token indices) of a merged run — tokenization_whisper.py:1286-1318, :1321-1350, :1353-1376, :1379-1405.
// Returns false when a token has no byte spelling.
bool split_words(const Vocab& v, const std::vector<int32_t>& ids, bool unspaced, std::vector<Word>& words) {
  const size_t n = ids.size();
  std::vector<uint8_t> all;
  std::vector<size_t> boff(n + 1, 0);
  for (size_t k = 0; k < n; ++k) {
    const int32_t t = ids[k];
    if (t < 0 || t >= v.eos_id || !v.has_bytes[t]) return false;
    all.insert(all.end(), v.bytes + v.off[t], v.bytes + v.off[t + 1]);
    boff[k + 1] = all.size();
  }
  ustr whole;
  utf8_decode_replace(all.data(), all.size(), whole);
  // smallest runs of tokens that decode to complete unicode
  struct Unit { ustr text; size_t first, last; };
  std::vector<Unit> units;
  size_t lo = 0, offset = 0;
  ustr s;
  for (size_t k = 0; k < n; ++k) {
    s.clear();
    utf8_decode_replace(all.data() + boff[lo], boff[k + 1] - boff[lo], s);
    const size_t p = s.find(kRepl);
    bool closed;
    if (p == ustr::npos) closed = true;
    else {
      if (offset + p >= whole.size()) return false;   // Python: IndexError
      closed = (whole[offset + p] == kRepl);
    }
    if (closed) {
      units.push_back(Unit{s, lo, k});
      lo = k + 1;
      offset += s.size();
    }
  }
  // words
  std::vector<ustr> wt;
  std::vector<std::vector<size_t>> wi;   // token indices (kept as lists: punctuation merging concatenates them)
  if (unspaced) {
    for (const Unit& u : units) {
      wt.push_back(u.text);
      std::vector<size_t> ix;
      for (size_t k = u.first; k <= u.last; ++k) ix.push_back(k);
      wi.push_back(ix);
    }
  } else {
    for (const Unit& u : units) {
      const bool opens = (ids[u.first] >= v.eos_id) || (!u.text.empty() && u.text[0] == U' ') ||
                         py_in(py_strip(u.text), kAsciiPunct) || wt.empty();
      if (opens) { wt.push_back(u.text); wi.emplace_back(); }
      else wt.back() += u.text;
      for (size_t k = u.first; k <= u.last; ++k) wi.back().push_back(k);
    }
  }
  // glue opening punctuation to the following word and closing punctuation to the preceding one
  const size_t m = wt.size();
  if (m >= 2) {
    size_t j = m - 1;
    for (size_t ii = m - 1; ii-- > 0;) {
      const ustr& w = wt[ii];
      if (!w.empty() && w[0] == U' ' && py_in(py_strip(w), kPrepended)) {
        wt[j] = w + wt[j];
        std::vector<size_t> ix = wi[ii];
        ix.insert(ix.end(), wi[j].begin(), wi[j].end());
        wi[j] = ix;
        wt[ii].clear(); wi[ii].clear();
      } else {
        j = ii;
      }
    }
    size_t i = 0;
    for (size_t jj = 1; jj < m; ++jj) {
      const bool ends_space = !wt[i].empty() && wt[i].back() == U' ';
      if (!ends_space && py_in(wt[jj], kAppended)) {
        wt[i] += wt[jj];
        wi[i].insert(wi[i].end(), wi[jj].begin(), wi[jj].end());
        wt[jj].clear(); wi[jj].clear();
      } else {
        i = jj;
      }
    }
  }
  // Python filters words, tokens and indices separately (`if w`, `if t`, `if x`): they stay aligned only when emptiness
  // coincides, which holds except for a unit of zero tokens (impossible) — an empty text with tokens means misalignment
  std::vector<ustr> fw;
  std::vector<std::vector<size_t>> fi;
  for (size_t k = 0; k < wt.size(); ++k) if (!wt[k].empty()) fw.push_back(wt[k]);
  for (size_t k = 0; k < wi.size(); ++k) if (!wi[k].empty()) fi.push_back(wi[k]);
  if (fw.size() != fi.size()) return false;   // Python's zip would silently truncate: let the Python path answer
  words.clear();
  for (size_t k = 0; k < fw.size(); ++k) words.push_back(Word{fw[k], fi[k].front(), fi[k].back()});
  return true;
}

struct Out {
  std::string text_bytes;              // raw bytes of every chunk's merged run, chunk after chunk
  std::vector<int64_t> chunk_off;      // [n_chunks + 1]
  std::string word_text;               // UTF-8 of every word
  std::vector<int64_t> word_off;       // [n_words + 1]
  std::vector<double> w_start, w_end;
  std::vector<int32_t> w_lang;
};

}  // namespace

extern "C" int cw_words_from_tokens(const uint8_t* tok_bytes, const int64_t* tok_off, const uint8_t* tok_has_bytes, int32_t eos_id,
                                    const uint8_t* is_special, const int32_t* lang_of, int32_t n_ids,
                                    const uint8_t* lang_unspaced, int32_t n_lang, int32_t default_unspaced,
                                    int32_t timestamp_begin, int32_t prompt_id, int32_t sot_id,
                                    int32_t n_outputs, const int32_t* tokens, const double* token_times, const int64_t* out_off,
                                    const double* strides, const uint8_t* has_stride, double time_precision, int32_t segment_size,
                                    char* text_buf, int64_t text_cap, int64_t* chunk_off, int32_t chunk_cap, int32_t* n_chunks,
                                    char* word_buf, int64_t word_cap, int64_t* word_off, double* word_start, double* word_end,
                                    int32_t* word_lang, int32_t words_cap, int32_t* n_words, int32_t* flags) {
  if (!tok_bytes || !tok_off || !tok_has_bytes || !is_special || !lang_of || !tokens || !token_times || !out_off || !strides ||
      !has_stride || !text_buf || !chunk_off || !n_chunks || !word_buf || !word_off || !word_start || !word_end || !word_lang ||
      !n_words || !flags || n_outputs < 0)
    return CW_ERR_INVALID;
  Vocab v{tok_bytes, tok_off, tok_has_bytes, eos_id, is_special, lang_of, n_ids, lang_unspaced, n_lang};
  const int32_t ts0 = timestamp_begin;
  Out o;
  o.chunk_off.push_back(0);
  o.word_off.push_back(0);
  *flags = 0;

  int32_t language = -1;                       // index into the caller's language list, -1 = None
  double chunk_t0 = 0.0; bool chunk_has_t0 = false;
  double time_offset = 0.0;
  std::vector<std::vector<int32_t>> held;
  std::vector<std::vector<Stamp>> held_st;
  bool skip = false;
  std::vector<int32_t> merged;
  std::vector<Stamp> merged_st;
  std::vector<Word> words;

  // chunk text + word chunks of the merged held runs (decode_asr.close / the final flush)
  auto emit_chunk = [&]() -> int {
    merge_overlaps(held, held_st, merged, merged_st);
    for (int32_t t : merged) {
      if (t < 0 || t >= v.eos_id || !v.has_bytes[t]) return CW_POST_PUNT;
      o.text_bytes.append((const char*)v.bytes + v.off[t], (size_t)(v.off[t + 1] - v.off[t]));
    }
    o.chunk_off.push_back((int64_t)o.text_bytes.size());
    const bool unspaced = language >= 0 ? (language < v.n_lang && v.lang_unspaced[language]) : (default_unspaced != 0);
    if (!split_words(v, merged, unspaced, words)) return CW_POST_PUNT;
    for (const Word& w : words) {
      if (w.first >= merged_st.size() || w.last >= merged_st.size()) return CW_POST_PUNT;   // Python: IndexError
      utf8_encode(w.text, o.word_text);
      o.word_off.push_back((int64_t)o.word_text.size());
      o.w_start.push_back(merged_st[w.first].a);
      o.w_end.push_back(merged_st[w.last].b);
      o.w_lang.push_back(language);
    }
    held.clear(); held_st.clear();
    chunk_has_t0 = false;
    return CW_OK;
  };

  for (int32_t oi = 0; oi < n_outputs; ++oi) {
    const int32_t* ids_all = tokens + out_off[oi];
    const double* times = token_times + out_off[oi];
    const int64_t n_all = out_off[oi + 1] - out_off[oi];
    int64_t beg = 0, n = n_all;
    if (n_all > 0 && ids_all[0] == prompt_id) {   // drop a <|startofprev|> prompt
      int64_t k = 0;
      while (k < n_all && ids_all[k] != sot_id) ++k;
      beg = k; n = n_all - k;                      // no sot: nothing left
    }
    const int32_t* ids = ids_all + beg;
    bool have_stride_end = false;
    int32_t stride_end_token = 0;
    double first_timestamp = (double)ts0;
    double seg_max = 0.0, seg_prev_max = 0.0, segs_before = 0.0;
    double chunk_len = 0.0, stride_right = 0.0;
    if (has_stride[oi]) {
      chunk_len = strides[3 * oi];
      const double stride_left = strides[3 * oi + 1];
      stride_right = strides[3 * oi + 2];
      time_offset -= stride_left;
      const double right_start = chunk_len - stride_right;
      if (stride_left != 0.0) first_timestamp = stride_left / time_precision + (double)ts0;
      if (stride_right != 0.0) {
        for (int64_t k = n; k-- > 0;) {
          const int32_t t = ids[k];
          if (t >= ts0) {
            if (have_stride_end && (double)(t - ts0) * time_precision < right_start) break;
            stride_end_token = t; have_stride_end = true;
          }
        }
      }
    }
    std::vector<int32_t> current;
    std::vector<Stamp> current_st;
    for (int64_t i = 0; i < n; ++i) {
      const int32_t t = ids[i];
      if (t < 0 || t >= n_ids) return CW_POST_PUNT;
      if (is_special[t]) {
        const int32_t name = lang_of[t];
        if (name >= 0) language = name;          // (the non-timestamp language-switch split does not apply in word mode)
      } else if (t >= ts0) {
        const double stamp = (double)(t - ts0) * time_precision;
        if (stamp < seg_max) {                   // timestamps restarted: generate() concatenated another 30 s segment
          const bool single_ending = i >= 2 && !(ids[i - 1] >= ts0 && ids[i - 2] >= ts0);
          if (single_ending) segs_before += time_precision * (double)segment_size;
          else { seg_max = seg_prev_max; segs_before += seg_prev_max; }
        }
        seg_prev_max = seg_max;
        seg_max = stamp;
        const double when = py_round2((double)(t - ts0) * time_precision + time_offset + segs_before);
        if (have_stride_end && stride_end_token != 0 && t >= stride_end_token) skip = true;
        else if (skip || (!held.empty() && (double)t < first_timestamp)) skip = false;
        else if (!chunk_has_t0) { chunk_t0 = when; chunk_has_t0 = true; }
        else if (when != chunk_t0) {
          held.push_back(current);
          held_st.push_back(current_st);
          const int rc = emit_chunk();
          if (rc != CW_OK) return rc;
          current.clear(); current_st.clear();
        }
      } else {
        current.push_back(t);
        // token_times is indexed by the position in the stripped ids, as in the reference
        if (i >= n_all) return CW_POST_PUNT;
        const double start = (i == 0) ? py_round2(0.0 + time_offset) : py_round2(times[i - 1] + time_offset);
        current_st.push_back(Stamp{start, py_round2(times[i] + time_offset)});
      }
    }
    if (has_stride[oi]) time_offset += chunk_len - stride_right;
    if (!current.empty()) {
      held.push_back(current);
      held_st.push_back(current_st);
    } else {
      bool any = false;
      for (const auto& h : held) any = any || !h.empty();
      if (!any) { chunk_has_t0 = false; held.clear(); held_st.clear(); }
    }
  }
  if (!held.empty()) {
    *flags |= 1;   // no closing timestamp token
    const int rc = emit_chunk();
    if (rc != CW_OK) return rc;
  }

  const int64_t nc = (int64_t)o.chunk_off.size() - 1, nw = (int64_t)o.word_off.size() - 1;
  if ((int64_t)o.text_bytes.size() > text_cap || nc > chunk_cap || (int64_t)o.word_text.size() > word_cap || nw > words_cap)
    return CW_ERR_WORKSPACE;
  memcpy(text_buf, o.text_bytes.data(), o.text_bytes.size());
  memcpy(chunk_off, o.chunk_off.data(), sizeof(int64_t) * (size_t)(nc + 1));
  memcpy(word_buf, o.word_text.data(), o.word_text.size());
  memcpy(word_off, o.word_off.data(), sizeof(int64_t) * (size_t)(nw + 1));
  if (nw) {
    memcpy(word_start, o.w_start.data(), sizeof(double) * (size_t)nw);
    memcpy(word_end, o.w_end.data(), sizeof(double) * (size_t)nw);
    memcpy(word_lang, o.w_lang.data(), sizeof(int32_t) * (size_t)nw);
  }
  *n_chunks = (int32_t)nc;
  *n_words = (int32_t)nw;
  return CW_OK;
}
