// main.cpp -- the `atm-raytracer` executable: CLI dispatch of the reference (main.rs:17-39) for the subcommands on the hot
// path -- `gen` and the three text dumpers that are its windows onto the stepper, the terrain sampler and the atmosphere
// (output-ray-paths, output-elev-profile, output-atm). `view` (an FLTK GUI) is outside the scope of this repository
// (SURVEY.md section 8).
#include <cstdio>
#include <cstring>

#include "host/atmrt_host.h"

int main(int argc, char** argv) {
    if (argc < 2 || strcmp(argv[1], "--help") == 0 || strcmp(argv[1], "help") == 0) {
        fprintf(stderr,
                "usage: atm-raytracer gen [-c CONFIG] [-t TERRAIN] [-l LAT] [-g LON] [-a ALT | -e ELEV] [-d DIR] [-f FOV]\n"
                "                         [-i TILT] [-m MAXDIST_KM] [--step M] [-R RADIUS_KM | --flat] [-s] [--output FILE]\n"
                "                         [--output-meta FILE] [-w WIDTH] [-h HEIGHT] [--gpus N]\n"
                "       atm-raytracer output-ray-paths INPUT [-h METERS] [-a MIN_DEG] [-b MAX_DEG] [-s DEG] [-r METERS] [-c METERS] [-o METERS]\n"
                "       atm-raytracer output-elev-profile INPUT [-a DEGREES] [-s METERS] [-c METERS]\n"
                "       atm-raytracer output-atm INPUT [-a ALTITUDE] [-b ALTITUDE] [-s LENGTH] [-c]\n");
        return argc < 2 ? 2 : 0;
    }
    const char* const* rest = argv + 2;
    if (strcmp(argv[1], "gen") == 0) return atmrt_host_gen(argc - 2, rest);
    if (strcmp(argv[1], "output-ray-paths") == 0) return atmrt_host_output_ray_paths(argc - 2, rest);
    if (strcmp(argv[1], "output-elev-profile") == 0) return atmrt_host_output_elev_profile(argc - 2, rest);
    if (strcmp(argv[1], "output-atm") == 0) return atmrt_host_output_atm(argc - 2, rest);
    fprintf(stderr, "ERROR: subcommand '%s' is not part of the B200 hot path (gen, output-ray-paths, output-elev-profile, output-atm are)\n", argv[1]);
    return 2;
}
