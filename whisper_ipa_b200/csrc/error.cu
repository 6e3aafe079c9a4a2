// Error plumbing and launch accounting of libwipa.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

int64_t g_wipa_launches = 0;
static int init_pdl() { const char* v = getenv("WIPA_PDL"); return (v && *v) ? atoi(v) : 0xff; }
int g_wipa_pdl = init_pdl();
static thread_local char g_wipa_err[1024] = "";

void wipa_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_wipa_err, sizeof(g_wipa_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* wipa_last_error(void) { return g_wipa_err; }

extern "C" const char* wipa_strerror(int code) {
    switch (code) {
        case WIPA_OK: return "ok";
        case WIPA_EINVAL: return "invalid argument";
        case WIPA_ECUDA: return "CUDA error";
        case WIPA_ENOMEM: return "out of memory";
        case WIPA_ESTATE: return "call out of order";
        case WIPA_EUNSUPPORTED: return "unsupported";
        default: return "unknown error";
    }
}

extern "C" int64_t wipa_launch_count(int reset) {
    const int64_t v = g_wipa_launches;
    if (reset) g_wipa_launches = 0;
    return v;
}
