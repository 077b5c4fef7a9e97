// prealpha_b200 -- command-line drop-in for `./prealpha general.inp` on the distribution path.
// Mirrors PROGRAM PREALPHA (Main:3843-3915): every argument is a general input file; without
// arguments "general.inp" is used (SETTINGS:73).  Options: --device N.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/prealpha_b200.h"

int main(int argc, char **argv)
{
    pa_exec ex;
    memset(&ex, 0, sizeof ex);
    ex.shard_count = 1;
    std::vector<std::string> inputs;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--device") && i + 1 < argc) ex.device = atoi(argv[++i]);
        else inputs.push_back(argv[i]);
    }
    if (inputs.empty()) inputs.push_back("general.inp");
    if (pa_device_count() < 1) {
        fprintf(stderr, "prealpha_b200: no usable B200 (sm_100) device; this program has no CPU fallback.\n");
        return 2;
    }
    int errors = 0;
    for (const std::string &g : inputs) errors += pa_run_general_input(g.c_str(), &ex);
    return errors ? 1 : 0;
}
