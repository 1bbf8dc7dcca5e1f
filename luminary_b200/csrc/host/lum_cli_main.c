/* lum_cli_main.c - headless command line front end: the benchmark mode of the reference's `LuminaryMD` without SDL.
 *
 *   LuminaryB200 [options] scene.lum [more.lum | mesh.obj ...]
 *     -b, --benchmark N NAME   queue the benchmark ladder of outputs up to 2^N samples and write
 *                              <out>/Bench-<samples>-<NAME>.png plus <out>/BenchResults-<NAME>.txt
 *     -o, --output DIR         output directory (default ".")
 *     --device ID              use only the given CUDA device(s); may be repeated (default: all)
 *     --supersampling S        (addition) override LuminaryRendererSettings.supersampling, which a version-4 scene file
 *                              cannot express; the library default is 1 = 2x2 internal resolution
 *     -v, --version / -h, --help
 *
 * Restates src/mandarin_duck/main.c:5-58, argument_parser.c:14-85 (the five options) and
 * mandarin_duck.c:53-98 (output ladder: 1,2,3,4,6,8,12,16,24,32, then every 32 samples from 64 to 2^N),
 * :186-244 (poll the promises, print "[time] N Samples", append "N, time" lines, save PNGs).
 * It only calls the public API of include/luminary/luminary.h. */
#include <luminary/luminary.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define CHECK(expr)                                                                                                    \
  do {                                                                                                                 \
    const LuminaryResult _r = (expr);                                                                                  \
    if (_r != LUMINARY_SUCCESS) {                                                                                      \
      fprintf(stderr, "%s failed: %s (%s)\n", #expr, luminary_result_to_string(_r), luminary_b200_last_error());       \
      return 1;                                                                                                        \
    }                                                                                                                  \
  } while (0)

static int ends_with(const char* s, const char* suffix) {
  const size_t n = strlen(s), m = strlen(suffix);
  return n >= m && strcmp(s + n - m, suffix) == 0;
}

int main(int argc, char** argv) {
  uint32_t num_benchmark_outputs = 0;
  const char* benchmark_name     = NULL;
  const char* output_directory   = ".";
  uint32_t device_mask           = LUMINARY_HOST_CREATE_INFO_DEVICE_MASK_ALL_DEVICES;
  int supersampling              = -1;
  int adaptive                   = -1; /* -1: keep what the scene file / defaults say */
  int adaptive_interval          = -1;
  const char* inputs[64];
  int num_inputs = 0;

  for (int i = 1; i < argc; i++) {
    const char* a = argv[i];
    if (!strcmp(a, "-b") || !strcmp(a, "--benchmark")) {
      if (i + 2 < argc + 0 && argv[i + 1][0] != '-') {
        num_benchmark_outputs = (uint32_t) atoll(argv[i + 1]);
        benchmark_name        = argv[i + 2];
        i += 2;
      }
    }
    else if (!strcmp(a, "-o") || !strcmp(a, "--output")) {
      if (i + 1 < argc)
        output_directory = argv[++i];
    }
    else if (!strcmp(a, "--device")) {
      if (i + 1 < argc) {
        if (device_mask == LUMINARY_HOST_CREATE_INFO_DEVICE_MASK_ALL_DEVICES)
          device_mask = 0;
        device_mask |= 1u << atoi(argv[++i]);
      }
    }
    else if (!strcmp(a, "--supersampling")) {
      if (i + 1 < argc)
        supersampling = atoi(argv[++i]);
    }
    else if (!strcmp(a, "--adaptive")) { /* 0 / 1: LuminaryRendererSettings.enable_adaptive_sampling */
      if (i + 1 < argc)
        adaptive = atoi(argv[++i]);
    }
    else if (!strcmp(a, "--adaptive-interval")) { /* executions of stage 0 before the first stage build (doubles per stage) */
      if (i + 1 < argc)
        adaptive_interval = atoi(argv[++i]);
    }
    else if (!strcmp(a, "-v") || !strcmp(a, "--version")) {
      printf("LuminaryB200 (B200-native path behind the Luminary host API)\n");
      return 0;
    }
    else if (!strcmp(a, "-h") || !strcmp(a, "--help")) {
      printf("USAGE: LuminaryB200 [options] file...\n\nOPTIONS:\n\t--benchmark, -b N NAME\n\t--output, -o DIR\n\t--device ID\n\t--supersampling S\n\t--adaptive 0|1\n\t--adaptive-interval N\n\t--version, -v\n\t--help, -h\n");
      return 0;
    }
    else if (a[0] == '-') {
      fprintf(stderr, "unknown option %s\n", a);
      return 1;
    }
    else if (num_inputs < 64)
      inputs[num_inputs++] = a;
  }
  if (num_inputs == 0 || num_benchmark_outputs == 0 || !benchmark_name) {
    fprintf(stderr, "nothing to do: this front end only implements the benchmark mode (-b N NAME) for at least one input file\n");
    return 1;
  }
  if (num_benchmark_outputs > 20) {
    fprintf(stderr, "at most 2^20 samples\n");
    return 1;
  }

  luminary_init();
  LuminaryHost* host;
  LuminaryHostCreateInfo info = {device_mask};
  CHECK(luminary_host_create(&host, info));

  for (int k = 0; k < num_inputs; k++) {
    LuminaryPath* path;
    CHECK(luminary_path_create(&path));
    CHECK(luminary_path_set_from_string(path, inputs[k]));
    if (ends_with(inputs[k], ".lum"))
      CHECK(luminary_host_load_lum_file(host, path));
    else if (ends_with(inputs[k], ".obj")) {
      CHECK(luminary_host_load_obj_file(host, path));
      LuminaryInstance inst;
      CHECK(luminary_host_new_instance(host, &inst));
      uint32_t meshes;
      CHECK(luminary_host_get_num_meshes(host, &meshes));
      inst.mesh_id = meshes - 1;
      CHECK(luminary_host_set_instance(host, &inst));
    }
    else {
      fprintf(stderr, "unsupported input file %s\n", inputs[k]);
      return 1;
    }
    CHECK(luminary_path_destroy(&path));
  }

  /* benchmark ladder, mandarin_duck.c:53-98 */
  LuminaryRendererSettings settings;
  CHECK(luminary_host_get_settings(host, &settings));
  if (supersampling >= 0)
    settings.supersampling = (uint32_t) supersampling;
  if (adaptive >= 0)
    settings.enable_adaptive_sampling = adaptive != 0;
  if (adaptive_interval > 0)
    settings.adaptive_sampling_update_interval = (uint32_t) adaptive_interval;
  if (supersampling >= 0 || adaptive >= 0 || adaptive_interval > 0)
    CHECK(luminary_host_set_settings(host, &settings));
  static LuminaryOutputPromiseHandle promises[40000];
  uint32_t num_promises = 0;
  const uint32_t num_exponential = (num_benchmark_outputs < 5) ? num_benchmark_outputs : 5;
  for (uint32_t k = 0; k <= num_exponential; k++) {
    LuminaryOutputRequestProperties p = {1u << k, settings.width, settings.height};
    CHECK(luminary_host_request_output(host, p, &promises[num_promises++]));
    if (k >= 2) {
      p.sample_count = (1u << (k - 1)) + (1u << (k - 2));
      CHECK(luminary_host_request_output(host, p, &promises[num_promises++]));
    }
  }
  if (num_benchmark_outputs > 5) {
    for (uint32_t s = 1u << 6; s <= (1u << num_benchmark_outputs) && num_promises < 40000; s += 32) {
      LuminaryOutputRequestProperties p = {s, settings.width, settings.height};
      CHECK(luminary_host_request_output(host, p, &promises[num_promises++]));
    }
  }

  char times_path[4096];
  snprintf(times_path, sizeof(times_path), "%s/BenchResults-%s.txt", output_directory, benchmark_name);
  FILE* times = fopen(times_path, "wb");
  if (!times) {
    fprintf(stderr, "Failed to open file %s\n", times_path);
    return 1;
  }

  CHECK(luminary_host_start_new_render(host));
  uint32_t obtained = 0;
  const struct timespec nap = {0, 2000000};
  while (obtained != num_promises) {
    for (uint32_t k = 0; k < num_promises; k++) {
      if (promises[k] == LUMINARY_OUTPUT_HANDLE_INVALID)
        continue;
      LuminaryOutputHandle out;
      CHECK(luminary_host_try_await_output(host, promises[k], &out));
      if (out == LUMINARY_OUTPUT_HANDLE_INVALID)
        continue;
      LuminaryImage image;
      CHECK(luminary_host_get_image(host, out, &image));
      printf("[%07.1fs] %05u Samples\n", image.meta_data.time, image.meta_data.sample_count);
      fflush(stdout);
      obtained++;
      char png[4096];
      snprintf(png, sizeof(png), "%s/Bench-%05u-%s.png", output_directory, image.meta_data.sample_count, benchmark_name);
      fprintf(times, "%u, %f\n", image.meta_data.sample_count, image.meta_data.time);
      LuminaryPath* path;
      CHECK(luminary_path_create(&path));
      CHECK(luminary_path_set_from_string(path, png));
      CHECK(luminary_host_save_png(host, out, path));
      CHECK(luminary_path_destroy(&path));
      CHECK(luminary_host_release_output(host, out));
      promises[k] = LUMINARY_OUTPUT_HANDLE_INVALID;
    }
    nanosleep(&nap, NULL);
  }
  fclose(times);

  uint64_t rays = 0;
  double seconds = 0.0;
  CHECK(luminary_b200_host_get_ray_count(host, &rays));
  CHECK(luminary_host_get_current_sample_time(host, &seconds));
  printf("%llu rays in %.3f GPU seconds: %.1f Mrays/s, %.2f samples/s\n", (unsigned long long) rays, seconds, rays / seconds * 1e-6,
         (double) (1u << num_benchmark_outputs) / seconds);

  CHECK(luminary_host_destroy(&host));
  luminary_shutdown();
  return 0;
}
