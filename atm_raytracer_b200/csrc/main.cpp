// main.cpp -- `atm-raytracer gen ...`: CLI dispatch of the reference (main.rs:17-39) for the one
// subcommand on the hot path. The other subcommands (view, output-atm, output-ray-paths,
// output-elev-profile) are outside the scope of this repository.
#include <cstdio>
#include <cstring>

extern "C" int atmrt_host_gen(int argc, const char* const* argv);

int main(int argc, char** argv) {
    if (argc < 2 || strcmp(argv[1], "--help") == 0 || strcmp(argv[1], "help") == 0) {
        fprintf(stderr,
                "usage: atm-raytracer gen [-c CONFIG] [-t TERRAIN] [-l LAT] [-g LON] [-a ALT | -e ELEV] [-d DIR] [-f FOV]\n"
                "                         [-i TILT] [-m MAXDIST_KM] [--step M] [-R RADIUS_KM | --flat] [-s] [--output FILE]\n"
                "                         [--output-meta FILE] [-w WIDTH] [-h HEIGHT]\n");
        return argc < 2 ? 2 : 0;
    }
    if (strcmp(argv[1], "gen") != 0) {
        fprintf(stderr, "ERROR: subcommand '%s' is not part of the B200 hot path (only `gen` is)\n", argv[1]);
        return 2;
    }
    return atmrt_host_gen(argc - 2, argv + 2);
}
