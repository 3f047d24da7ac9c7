// Development tool (tools/loader_fuzz/mutate.py): loads every file named on the command line through the library's readers;
// built with -fsanitize=address,undefined.
#include <cstdio>
#include <cstring>
#include <string>
#include "../../include/ygl_b200.h"
#include "../../yocto-gl_b200/csrc/ygl_hostio.h"
extern "C" void ygl_internal_set_error(const char*) {}
int main(int argc, char** argv) {
  int ok = 0, bad = 0;
  for (int k = 1; k < argc; k++) {
    std::string path = argv[k], ext = ygl_io::path_extension(path);
    if (ext == ".png" || ext == ".jpg" || ext == ".hdr" || ext == ".exr") {
      ygl_io::HostTexture tex; std::string err;
      (ygl_io::load_texture(path, tex, err) ? ok : bad)++;
    } else {
      ygl_loaded_scene* s = nullptr;
      if (ygl_scene_load(path.c_str(), &s) == 0) { ok++; ygl_loaded_scene_destroy(s); } else bad++;
    }
  }
  printf("loaded %d refused %d\n", ok, bad);
}
