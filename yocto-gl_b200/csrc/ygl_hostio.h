// ygl_hostio.h — what the host-side readers of the library share (scene files in ygl_sceneio.cpp, image files in
// ygl_imageio.cpp). Internal: nothing here is part of the C ABI.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

namespace ygl_io {

struct HostTexture {
  int                  width = 0, height = 0, linear = 0, nearest = 0, clamp = 0;
  std::vector<float>   pixelsf;  // rgba
  std::vector<uint8_t> pixelsb;  // rgba
};

// load_texture, yocto_sceneio.cpp:1796-1837: .png / .jpg -> byte rgba as stb_image returns them, .hdr / .exr -> float rgba as
// stb_image / tinyexr do; the file type decides `linear`
bool load_texture(const std::string& filename, HostTexture& tex, std::string& error);

bool        read_file(const std::string& filename, std::vector<uint8_t>& data, std::string& error);
std::string path_extension(const std::string& path);  // lower case, with the dot

}  // namespace ygl_io
