// version / error plumbing of libcolddiff
#include "cd_common.cuh"
#include <string.h>

static thread_local char g_err[512] = {0};

void cd_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int cd_version(void) { return CD_ABI_VERSION; }

extern "C" int cd_last_error(char* buf, size_t n) {
  const size_t len = strlen(g_err);
  if (buf && n) {
    const size_t c = len < n - 1 ? len : n - 1;
    memcpy(buf, g_err, c);
    buf[c] = 0;
  }
  return (int)len;
}
