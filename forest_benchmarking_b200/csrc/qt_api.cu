// Error plumbing and version for libqtomo.
#include "qt_common.cuh"
#include "../../include/qtomo.h"

#include <stdarg.h>
#include <string.h>

static thread_local char g_err[512] = "";

void qt_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int qt_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    qt_set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return QT_ERR_CUDA;
  }
  return QT_OK;
}

extern "C" int qt_version(void) { return 100; }

extern "C" int qt_last_error(char* buf, int len) {
  if (!buf || len <= 0) return QT_ERR_ARG;
  strncpy(buf, g_err, (size_t)len - 1);
  buf[len - 1] = '\0';
  return QT_OK;
}
