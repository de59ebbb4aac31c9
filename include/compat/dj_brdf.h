/* include/compat/dj_brdf.h -- the reference's header NAME, backed by the B200 engine.
 *
 * A program written against jdupuy/dj_brdf (`#define DJ_BRDF_IMPLEMENTATION 1` + `#include "dj_brdf.h"`)
 * builds unchanged with `-I<repo>/include/compat -L<repo>/dj_brdf_b200 -ldjb200`: the `djb::` classes it uses
 * are provided by djb200_facade.hpp and run on the GPU through the C-ABI.
 * DJ_BRDF_IMPLEMENTATION is accepted and ignored: there is no header-side implementation to emit. */
#ifndef DJB200_COMPAT_DJ_BRDF_H
#define DJB200_COMPAT_DJ_BRDF_H
#include "../djb200_facade.hpp"
#endif
