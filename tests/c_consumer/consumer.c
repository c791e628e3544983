/* A plain-C (C11) consumer of include/llamab200.h: the way a cgo shim under pkg/llama binds the
 * library (cgo compiles its preamble as C, not C++).  Mirrors the reference call sequence
 *   llama.LoadModel -> llama.NewContext -> llama.Eval -> lctx.Logits
 * (pkg/llama/llama.go:91-113, 211-218, 394-401).
 *
 *   consumer            : links, prints the version, exits 0 without a GPU ("NO_GPU")
 *   consumer <seed>     : with a GPU — create -> set_tensor -> eval(prompt) -> eval(1 token), prints the
 *                         logits of both calls as hex floats, one per line, for the Python test to compare
 *                         with the ctypes path.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "llamab200.h"

static int fail(const char *what) {
    fprintf(stderr, "consumer: %s: %s\n", what, lb_last_error());
    return 1;
}

int main(int argc, char **argv) {
    printf("version %s\n", lb_version());
    if (lb_device_count() <= 0) {
        printf("NO_GPU\n");
        return 0;
    }
    const uint64_t seed = argc > 1 ? strtoull(argv[1], NULL, 10) : 7;
    const lb_hparams hp = {.vocab = 96, .dim = 64, .mult = 32, .heads = 2, .layers = 2};
    lb_model *m = lb_model_create(&hp, 0, 0, hp.layers, LB_TYPE_F32);
    if (!m) return fail("lb_model_create");
    if (lb_model_init_random(m, seed)) return fail("lb_model_init_random");
    float norm[64];
    for (int i = 0; i < 64; i++) norm[i] = 1.0f + (float)i / 64.0f;   /* exactly representable: the Python side builds the same bits */
    if (lb_model_set_tensor(m, "norm.weight", LB_TYPE_F32, norm, sizeof norm)) return fail("lb_model_set_tensor");
    if (!lb_model_set_tensor(m, "layers.0.bogus.weight", LB_TYPE_F32, norm, sizeof norm)) {
        fprintf(stderr, "consumer: unknown tensor name was accepted\n");  /* llama.go:906-910 aborts */
        return 1;
    }
    lb_context *c = lb_context_create(m, 32);
    if (!c) return fail("lb_context_create");
    const uint32_t prompt[5] = {1, 35, 36, 90, 7}, next = 11;
    float *logits = malloc(sizeof(float) * hp.vocab);
    if (!logits) return 1;
    if (lb_eval(c, prompt, 5, 0, logits)) return fail("lb_eval(prompt)");
    for (uint32_t i = 0; i < hp.vocab; i++) printf("P %a\n", (double)logits[i]);
    if (lb_eval(c, &next, 1, 5, logits)) return fail("lb_eval(decode)");
    for (uint32_t i = 0; i < hp.vocab; i++) printf("D %a\n", (double)logits[i]);
    if (!lb_eval(c, prompt, 5, 30, logits)) {   /* pastCount + N > context must fail, not scribble */
        fprintf(stderr, "consumer: eval past the context was accepted\n");
        return 1;
    }
    free(logits);
    lb_context_free(c);
    lb_model_free(m);
    printf("OK\n");
    return 0;
}
