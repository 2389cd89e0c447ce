// wrh_image.cpp — Image.initFromFile of the host mirror (src/image.zig:12-17 -> zstbi.Image.loadFromFile,
// libs/zstbi/src/zstbi.zig:77): JPEG / PNG decode with the reference's OWN decoder, the stb_image v2.28 it vendors under
// libs/zstbi/libs/stbi.  The header is compiled WHERE IT LIES in the reference checkout (build.py adds
// -DWRH_HAVE_STBI -I<reference>/libs/zstbi/libs/stbi when it finds it, or -I$WRT_STBI_INCLUDE); no copy of it lives in
// this repository.  A Zig host keeps using zstbi itself and hands the decoded bytes over the C ABI (wrt_scene.texels).
#include <cstring>

#include "wrh_scene.hpp"

#ifdef WRH_HAVE_STBI
#define STB_IMAGE_IMPLEMENTATION
#define STB_IMAGE_STATIC
#if defined(__GNUC__)
#pragma GCC diagnostic push
#pragma GCC diagnostic ignored "-Wunused-function"
#pragma GCC diagnostic ignored "-Wsign-compare"
#pragma GCC diagnostic ignored "-Wmissing-field-initializers"
#pragma GCC diagnostic ignored "-Wunused-but-set-variable"
#endif
#include "stb_image.h"
#if defined(__GNUC__)
#pragma GCC diagnostic pop
#endif
#endif

namespace wrh {

bool Image::decoderAvailable() {
#ifdef WRH_HAVE_STBI
    return true;
#else
    return false;
#endif
}

bool Image::loadFromFile(const std::string& path, Image& out, std::string& why) {
#ifdef WRH_HAVE_STBI
    int w = 0, h = 0, comps = 0;
    unsigned char* px = stbi_load(path.c_str(), &w, &h, &comps, 0);  // forced_num_components = 0 (image.zig:15-16)
    if (!px) {
        why = std::string("stb_image: ") + (stbi_failure_reason() ? stbi_failure_reason() : "cannot decode") + ": " + path;
        return false;
    }
    out = Image::fromPixels(static_cast<uint32_t>(w), static_cast<uint32_t>(h), static_cast<uint32_t>(comps), px);
    stbi_image_free(px);
    return true;
#else
    (void)out;
    why = "this build has no JPEG/PNG decoder (the reference's stb_image.h was not found at build time; set WRT_STBI_INCLUDE): " + path;
    return false;
#endif
}

}  // namespace wrh
