/* ORACLE (test infrastructure): compiles the reference's vendored stb_image.h (public domain, v2.28) where it
 * lies under /root/reference/libs/zstbi/libs/stbi — the decoder zstbi.Image.loadFromFile wraps
 * (libs/zstbi/src/zstbi.zig:77, src/image.zig:12-17).  Used to decode the reference assets into golden texel
 * fixtures exactly as the reference would see them.  No reference source is copied into this repository. */
#define STB_IMAGE_IMPLEMENTATION
#define STBI_NO_STDIO_UNUSED
#include "stb_image.h"

unsigned char* wro_stbi_load(const char* path, int* w, int* h, int* comps) {
    return stbi_load(path, w, h, comps, 0); /* forced_num_components = 0, image.zig:15-16 */
}
void wro_stbi_free(unsigned char* p) { stbi_image_free(p); }
