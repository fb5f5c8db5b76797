// shim_common.hpp - shared by the drop-in shims: turns a C-ABI status into the reference's
// "void, no error channel" contract (SURVEY 8b "Errors"): message on stderr, then abort().
#pragma once
#include <cstdio>
#include <cstdlib>
#include "../../include/pfdr_b200.h"

static inline void pfdr_b200_shim_check(int status, const char *who)
{
    if (status != PFDR_B200_OK) {
        fprintf(stderr, "%s: libpfdr_b200 failed with status %d: %s\n", who, status, pfdr_b200_last_error());
        abort();
    }
}
