/* example_render.c - the whole path through the C ABI in plain C: read a scene file (Yocto/GL JSON, .ply, .obj, glTF),
 * render it on the GPU, tonemap on the GPU, write a binary PPM. What apps/ytrace.cpp:41-160 does with the reference
 * library, minus the command line.
 *   gcc -std=c99 -I include yocto-gl_b200/host/example_render.c -o example_render \
 *       -L yocto-gl_b200/lib -l:libygl_b200.so -Wl,-rpath,$PWD/yocto-gl_b200/lib
 *   ./example_render scene.json out.ppm [resolution] [samples]
 * There is no CPU path: without a CUDA device the program reports the library's error and exits with status 2. */
#include <stdio.h>
#include <stdlib.h>

#include "ygl_b200.h"

static int fail(const char* what) {
  fprintf(stderr, "%s: %s\n", what, ygl_last_error());
  return 2;
}

int main(int argc, char** argv) {
  if (argc < 3) return fprintf(stderr, "usage: %s <scene file> <out.ppm> [resolution] [samples]\n", argv[0]), 1;
  ygl_loaded_scene* loaded = NULL;
  if (ygl_scene_load(argv[1], &loaded) != YGL_OK) return fail("ygl_scene_load");
  const ygl_scene_desc* desc = ygl_loaded_scene_desc(loaded);
  printf("%s: %d shapes, %d instances, %d materials, %d textures, %d cameras, %d environments\n", argv[1], desc->num_shapes,
      desc->num_instances, desc->num_materials, desc->num_textures, desc->num_cameras, desc->num_environments);

  ygl_trace_params params;
  ygl_trace_params_default(&params);
  if (argc > 3) params.resolution = atoi(argv[3]);
  if (argc > 4) params.samples = atoi(argv[4]);

  ygl_context* ctx = NULL;
  if (ygl_context_create(0, &ctx) != YGL_OK) {
    ygl_loaded_scene_destroy(loaded);
    return fail("ygl_context_create");
  }
  int width = 0, height = 0, status = 0;
  if (ygl_trace_image(ctx, desc, &params, &width, &height, NULL) != YGL_OK) status = fail("ygl_trace_image (size)");
  float*         hdr = status ? NULL : (float*)malloc((size_t)width * height * 4 * sizeof(float));
  unsigned char* ldr = status ? NULL : (unsigned char*)malloc((size_t)width * height * 4);
  if (!status && ygl_trace_image(ctx, desc, &params, &width, &height, hdr) != YGL_OK) status = fail("ygl_trace_image");
  if (!status && ygl_tonemap_image(ctx, hdr, (int64_t)width * height, 0.0f, /*filmic*/ 0, /*srgb*/ 1, NULL, ldr) != YGL_OK)
    status = fail("ygl_tonemap_image");
  if (!status) {
    FILE* f = fopen(argv[2], "wb");
    if (!f) {
      perror(argv[2]);
      status = 1;
    } else {
      fprintf(f, "P6\n%d %d\n255\n", width, height);
      for (long i = 0; i < (long)width * height; i++) fwrite(ldr + 4 * i, 1, 3, f);
      fclose(f);
      printf("%s: %d x %d, %d samples per pixel\n", argv[2], width, height, params.samples);
    }
  }
  free(hdr), free(ldr);
  ygl_context_destroy(ctx);
  ygl_loaded_scene_destroy(loaded);
  return status;
}
