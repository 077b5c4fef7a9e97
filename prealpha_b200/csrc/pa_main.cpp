// prealpha_b200 -- command-line drop-in for `./prealpha general.inp` on the distribution path.
// Mirrors PROGRAM PREALPHA (Main:3843-3915): every argument is a general input file; without
// arguments "general.inp" is used (SETTINGS:73).  Options: --device N (first device), --devices N | -g N | -g all
// (how many GPUs one analysis may use; default: all usable devices for large jobs, see pa_exec.ndevices;
// environment PREALPHA_B200_DEVICES has the same meaning).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
#include <chrono>
#include <string>
#include <vector>
#include "../../include/prealpha_b200.h"

int main(int argc, char **argv)
{
    pa_exec ex;
    memset(&ex, 0, sizeof ex);
    ex.shard_count = 1;
    ex.tile_shard_count = 1;
    ex.ndevices = -1;                                            // all usable devices, for jobs that are worth it
    auto parse_devices = [](const char *v) { return (!strcmp(v, "all") || !strcmp(v, "0")) ? -1 : atoi(v); };
    if (const char *e = getenv("PREALPHA_B200_DEVICES")) ex.ndevices = parse_devices(e);
    std::vector<std::string> inputs;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--device") && i + 1 < argc) ex.device = atoi(argv[++i]);
        else if ((!strcmp(argv[i], "--devices") || !strcmp(argv[i], "-g")) && i + 1 < argc) ex.ndevices = parse_devices(argv[++i]);
        else inputs.push_back(argv[i]);
    }
    if (inputs.empty()) inputs.push_back("general.inp");
    // An explicit device count: make only those devices visible to the CUDA driver.  Without a persistence daemon the
    // driver initialises every visible GPU at start-up (about a second each: 8.8 s on an 8-GPU box), which a run on one
    // GPU need not pay.  (The default, "all usable devices for jobs that are worth it", has to see them all.)
    if (ex.ndevices >= 1 && !getenv("CUDA_VISIBLE_DEVICES")) {
        std::string vis;
        for (int d = 0; d < ex.ndevices; ++d) vis += (d ? "," : "") + std::to_string(ex.device + d);
        setenv("CUDA_VISIBLE_DEVICES", vis.c_str(), 1);
        ex.device = 0;
    }
    const bool dbg_t = getenv("PA_DEBUG_TIMING") != nullptr;
    const auto t_main = std::chrono::steady_clock::now();
    auto since_main = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_main).count(); };
    const int usable = pa_device_count();                        // first CUDA call: the driver initialises here
    if (dbg_t) fprintf(stderr, "[prealpha_b200] main: driver initialised, %d usable device(s) %9.1f ms\n", usable, since_main());
    if (usable < 1) {
        fprintf(stderr, "prealpha_b200: no usable B200 (sm_100) device; this program has no CPU fallback.\n");
        return 2;
    }
    // Exit status: 0 like the reference binary, whose warnings and errors only go to the report ("Encountered N
    // errors/warnings globally"); 3 when a GPU entry point failed (pa_run_general_input < 0), 2 without a usable device.
    int failures = 0;
    for (const std::string &g : inputs) if (pa_run_general_input(g.c_str(), &ex) < 0) ++failures;
    fflush(stdout);
    if (dbg_t) fprintf(stderr, "[prealpha_b200] main: leaving %9.1f ms\n", since_main());
    // the results are on disk: skip the orderly teardown of CUDA contexts, memory pools and NCCL communicators, which
    // costs up to a second per device and frees nothing the operating system does not free anyway
    _exit(failures ? 3 : 0);
}
