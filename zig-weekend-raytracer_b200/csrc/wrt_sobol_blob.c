/* wrt_sobol_blob.c — embeds the Sobol / van-der-Corput tables (data extracted from the reference's
 * src/math/sobolmatrices.zig by tools/gen_sobol_tables.py; layout documented there) into libwrt.so. */
#include <stdint.h>

#ifndef WRT_SOBOL_BLOB
#error "define WRT_SOBOL_BLOB to the absolute path of data/sobol_tables.bin"
#endif

__asm__(".section .rodata\n"
        ".balign 16\n"
        ".global wrt_sobol_blob\n"
        ".hidden wrt_sobol_blob\n"
        "wrt_sobol_blob:\n"
        ".incbin \"" WRT_SOBOL_BLOB "\"\n"
        ".global wrt_sobol_blob_end\n"
        ".hidden wrt_sobol_blob_end\n"
        "wrt_sobol_blob_end:\n"
        ".previous\n");
