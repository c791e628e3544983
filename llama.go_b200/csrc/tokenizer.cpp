// tokenizer.cpp — the reference's SentencePiece-style tokenizer on the host.
//
// Same algorithm as pkg/ml/ml.go:2761-2848 (itself llama.cpp's llama_tokenize): split the text into
// UTF-8 characters by the first byte's length class (utf8Len, ml.go:2705-2709), seed a work queue with
// every adjacent pair that concatenates to a vocab entry, repeatedly merge the pair with the highest
// score (ties: the leftmost pair, PopMax ml.go:2719-2737), re-offer the new neighbours, and finally emit
// each surviving symbol's id — or, when the symbol is not in the vocab, its bytes as ids byte + 3
// (ml.go:2827-2833).  Pure host code: usable without a GPU.
#include <stdint.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace lb {
namespace tok {

struct Vocab {  // ml.Vocab (ml.go:2653-2657)
    std::vector<std::string> id2token;
    std::vector<float> score;
    std::unordered_map<std::string, uint32_t> token2id;
};

static uint32_t utf8_len(unsigned char c) {
    static const uint32_t lookup[16] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 3, 4};
    return lookup[c >> 4];
}

struct Symbol { int prev, next; size_t off; uint32_t n; };
struct Bigram { int left, right; float score; uint32_t size; };

std::vector<uint32_t> tokenize(const Vocab &v, const std::string &text, bool bos) {
    std::vector<uint32_t> out;
    if (bos) out.push_back(1);  // TOKEN_BOS, ml.go:26,2767-2769
    if (text.empty()) return out;  // (the reference indexes symbols[0] unconditionally; an empty text never reaches it)
    std::vector<Symbol> sym;
    for (size_t off = 0; off < text.size();) {
        uint32_t len = utf8_len((unsigned char)text[off]);
        if (len > text.size() - off) len = (uint32_t)(text.size() - off);
        Symbol s;
        s.off = off; s.n = len;
        s.prev = (int)sym.size() - 1;
        off += len;
        s.next = off == text.size() ? -1 : (int)sym.size() + 1;
        sym.push_back(s);
    }
    std::vector<Bigram> queue;
    auto try_add = [&](int left, int right) {  // TryAddBigram, ml.go:2739-2756
        if (left == -1 || right == -1) return;
        const std::string piece = text.substr(sym[left].off, sym[left].n + sym[right].n);
        auto it = v.token2id.find(piece);
        if (it == v.token2id.end() || it->second >= v.id2token.size()) return;
        queue.push_back({left, right, v.score[it->second], (uint32_t)piece.size()});
    };
    for (size_t i = 1; i < sym.size(); i++) try_add((int)i - 1, (int)i);
    while (!queue.empty()) {
        size_t mx = 0;  // PopMax: highest score, ties -> smaller left index
        for (size_t cur = 1; cur < queue.size(); cur++)
            if (queue[mx].score < queue[cur].score || (queue[mx].score == queue[cur].score && queue[mx].left > queue[cur].left)) mx = cur;
        const Bigram b = queue[mx];
        queue[mx] = queue.back();
        queue.pop_back();
        Symbol &l = sym[b.left], &r = sym[b.right];
        if (l.n == 0 || r.n == 0 || l.n + r.n != b.size) continue;  // one side was merged already
        l.n += r.n;
        r.n = 0;
        l.next = r.next;
        if (r.next >= 0) sym[r.next].prev = b.left;
        try_add(l.prev, b.left);
        try_add(b.left, l.next);
    }
    for (int i = 0; i != -1; i = sym[i].next) {
        const Symbol &s = sym[i];
        auto it = v.token2id.find(text.substr(s.off, s.n));
        if (it == v.token2id.end()) {
            for (uint32_t j = 0; j < s.n; j++) out.push_back((uint32_t)(uint8_t)((unsigned char)text[s.off + j] + 3));  // byte fallback in Go `byte` arithmetic: 0xFD..0xFF wrap to 0..2 (ml.go:2831)
        } else {
            out.push_back(it->second);
        }
    }
    return out;
}

}  // namespace tok
}  // namespace lb

// ---- C-ABI ----
#include "../../include/llamab200.h"

struct lb_vocab { lb::tok::Vocab v; };
extern thread_local std::string g_lb_tok_err;
thread_local std::string g_lb_tok_err;

extern "C" {
lb_vocab *lb_vocab_create(uint32_t size) {
    auto *h = new lb_vocab();
    h->v.id2token.resize(size);
    h->v.score.resize(size, 0.f);
    return h;
}
void lb_vocab_free(lb_vocab *v) { delete v; }
int lb_vocab_set(lb_vocab *v, uint32_t id, const char *bytes, uint32_t len, float score) {
    if (!v || id >= v->v.id2token.size() || (!bytes && len)) return 1;
    v->v.id2token[id].assign(bytes ? bytes : "", len);
    v->v.score[id] = score;
    v->v.token2id[v->v.id2token[id]] = id;  // later ids win, like the loader's map assignment (llama.go:809)
    return 0;
}
int64_t lb_tokenize(const lb_vocab *v, const char *text, uint32_t len, int bos, uint32_t *out, uint32_t cap) {
    if (!v || (!text && len)) return -1;
    std::vector<uint32_t> ids = lb::tok::tokenize(v->v, std::string(text ? text : "", len), bos != 0);
    if (out)
        for (size_t i = 0; i < ids.size() && i < cap; i++) out[i] = ids[i];
    return (int64_t)ids.size();
}
}
