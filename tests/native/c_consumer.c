/* A plain C (C99) consumer of include/b200exec.h: what a cgo / JNI / Rust `extern "C"` binding sees.  Built and run by
 * tests/test_abi.py without a GPU: only the host-only entry points are called (version, protobuf plan decoder, typed plan,
 * TaskStatus encoder).  usage: c_consumer <file with PhysicalPlanNode bytes> */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200exec.h"

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  static unsigned char buf[1 << 20];
  size_t n = fread(buf, 1, sizeof buf, f);
  fclose(f);

  printf("version=%s\n", b200_version());
  printf("sizeof_task_result=%zu sizeof_swp=%zu sizeof_metrics=%zu\n", sizeof(b200_task_result), sizeof(b200_shuffle_write_partition),
         sizeof(b200_operator_metrics));

  char* json = NULL;
  int rc = b200_plan_proto_to_json(buf, (uint64_t)n, "job-from-c", &json);
  if (rc != B200_OK) {
    printf("decode failed: %d %s\n", rc, b200_last_error());
    return 1;
  }
  char* typed = NULL;
  rc = b200_plan_typed_json(json, 0, &typed);
  if (rc != B200_OK) {
    printf("typing failed: %d %s\n", rc, b200_last_error());
    return 1;
  }
  printf("ir_bytes=%zu typed_bytes=%zu has_job=%d\n", strlen(json), strlen(typed), strstr(json, "job-from-c") != NULL);
  b200_string_free(json);
  b200_string_free(typed);

  /* a malformed message is an error code, not a crash */
  rc = b200_plan_proto_to_json("\x0a\xff\xff\xff\xff\x0f", 6, NULL, &json);
  printf("malformed_rc=%d\n", rc);

  b200_task_result r;
  memset(&r, 0, sizeof r);
  r.task_id = 17;
  r.stage_id = 5;
  r.partition_id = 3;
  r.status = B200_OK;
  b200_shuffle_write_partition p;
  memset(&p, 0, sizeof p);
  p.partition_id = 1;
  p.num_rows = 10;
  p.num_batches = 1;
  p.num_bytes = 160;
  p.file_id = -1;
  char* status = NULL;
  uint64_t status_len = 0;
  rc = b200_task_status_encode("job-from-c", "exec-c", &r, &p, 1, NULL, 0, &status, &status_len);
  printf("status_rc=%d status_len=%llu\n", rc, (unsigned long long)status_len);
  b200_string_free(status);
  return 0;
}
