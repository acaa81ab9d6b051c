/* C harness for the drop-in boundary: includes include/mindrec_b200.h, loads libmindrec_b200.so with dlopen (no
 * Python, no torch) and drives the argument validation of a few aot entry points plus the plain-C helpers.
 * Nothing here launches a kernel, so it runs on a machine without a GPU (tests/test_abi.py builds and runs it).
 *
 *   gcc -std=c99 -I include tests/c/abi_harness.c -ldl -o abi_harness && ./abi_harness mindrec_b200/libmindrec_b200.so
 */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "mindrec_b200.h"

typedef int (*aot_fn)(int, void **, int *, int64_t **, const char **, void *, void *);
typedef const char *(*str_fn)(void);
typedef size_t (*ws2_fn)(int64_t, int);
typedef size_t (*ws1_fn)(int64_t);

static int failures = 0;
#define CHECK(cond, what)                                        \
  do {                                                           \
    if (!(cond)) { printf("FAIL: %s\n", what); ++failures; }     \
    else printf("ok:   %s\n", what);                             \
  } while (0)

int main(int argc, char **argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s <path to libmindrec_b200.so>\n", argv[0]); return 2; }
  void *h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
  str_fn version = (str_fn)dlsym(h, "mrec_version");
  str_fn last_error = (str_fn)dlsym(h, "mrec_last_error");
  aot_fn gather = (aot_fn)dlsym(h, "mrec_gather");
  aot_fn lazy_adam = (aot_fn)dlsym(h, "mrec_sparse_lazy_adam");
  ws2_fn unique_ws = (ws2_fn)dlsym(h, "mrec_unique_workspace_bytes");
  ws1_fn relu_ws = (ws1_fn)dlsym(h, "mrec_relu_bwd_bias_workspace_bytes");
  CHECK(version && last_error && gather && lazy_adam && unique_ws && relu_ws, "symbols resolve");
  if (failures) return 1;
  CHECK(strstr(version(), "sm_100a") != NULL, "mrec_version names sm_100a");

  /* mrec_gather(table[4,4] f32, ids[2] i32 -> out[2,4] f32): validation only, the pointers are never dereferenced */
  int64_t s_table[2] = {4, 4}, s_ids[1] = {2}, s_out[2] = {2, 4}, s_bad[2] = {3, 4};
  int64_t *shapes[3] = {s_table, s_ids, s_out};
  int ndims[3] = {2, 1, 2};
  const char *dtypes[3] = {"float32", "int32", "float32"};
  void *params[3] = {(void *)16, (void *)32, (void *)64};
  CHECK(gather(2, params, ndims, shapes, dtypes, NULL, NULL) == 1, "wrong nparam -> MREC_ERR_NPARAM (1)");
  CHECK(strlen(last_error()) > 0, "mrec_last_error explains it");
  const char *bad_dt[3] = {"float16", "int32", "float32"};
  CHECK(gather(3, params, ndims, shapes, bad_dt, NULL, NULL) == 2, "float16 table -> MREC_ERR_DTYPE (2)");
  int64_t *bad_shapes[3] = {s_table, s_ids, s_bad};
  CHECK(gather(3, params, ndims, bad_shapes, dtypes, NULL, NULL) == 3, "out[3,4] for 2 ids -> MREC_ERR_SHAPE (3)");
  void *misaligned[3] = {(void *)20, (void *)32, (void *)64};
  CHECK(gather(3, misaligned, ndims, shapes, dtypes, NULL, NULL) == 4, "table at a 4-byte boundary -> MREC_ERR_ALIGN (4)");
  void *nulls[3] = {NULL, (void *)32, (void *)64};
  CHECK(gather(3, nulls, ndims, shapes, dtypes, NULL, NULL) == 8, "null table -> MREC_ERR_NULL (8)");
  CHECK(lazy_adam(0, NULL, NULL, NULL, NULL, NULL, NULL) == 1, "mrec_sparse_lazy_adam with no params -> 1");

  CHECK(unique_ws(624000, 4) >= (size_t)624000 * 12, "mrec_unique_workspace_bytes(624000, 4) covers two key buffers + values");
  CHECK(relu_ws(1024) >= (size_t)1024 * 4, "mrec_relu_bwd_bias_workspace_bytes(1024)");
  dlclose(h);
  printf("%s\n", failures ? "ABI HARNESS FAILED" : "ABI HARNESS OK");
  return failures ? 1 : 0;
}
