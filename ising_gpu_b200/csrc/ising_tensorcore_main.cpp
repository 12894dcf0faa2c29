// ising_tensorcore_main.cpp -- command-line host of the tensor-core variant with the CLI surface of
// the reference's tensorcore/main.cu (:402-451 options, :580-609 warm-up/trial/report, :338-348
// final.txt).  Pure C++ over the C ABI (ising_tc_* in include/ising_b200.h).
#include <getopt.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <iostream>

#include "ising_b200.h"

#define TC_CALL(call)                                                          \
    do {                                                                       \
        if ((call) != ISING_OK) {                                              \
            fprintf(stderr, "%s:%d: %s\n", __FILE__, __LINE__, ising_last_error()); \
            exit(EXIT_FAILURE);                                                \
        }                                                                      \
    } while (0)

static void usage(const char* pname)
{
    const char* bname = strrchr(pname, '/');
    bname = bname ? bname + 1 : pname;
    fprintf(stdout,
            "Usage: %s [options]\n"
            "options:\n"
            "\t-x|--lattice-nbx <LATTICE_NBX>\n"
            "\t\tnumber of blocks along lattice rows (number of rows / 256)\n\n"
            "\t-y|--lattice-nby <LATTICE_NBY>\n"
            "\t\tnumber of blocks along lattice columns (number of columns / 256)\n\n"
            "\t-g|--ngpus <NGPUS>\n"
            "\t\tnumber of GPUs to use for simulation\n\n"
            "\t-w|--nwarmup <NWARMUP>\n"
            "\t\tnumber of warmup iterations\n\n"
            "\t-n|--niters <NITERS>\n"
            "\t\tnumber of trial iterations\n\n"
            "\t-a|--alpha <ALPHA>\n"
            "\t\tcoefficient of critical temperature\n\n"
            "\t-s|--seed <SEED>\n"
            "\t\tseed for random number generation\n\n"
            "\t-o|--write-lattice\n"
            "\t\twrite final lattice configuration to file\n\n",
            bname);
    exit(EXIT_SUCCESS);
}

int main(int argc, char** argv)
{
    int nbx = 10, nby = 10, niter = 1000, nwarmup = 100, nGPUs = 1;   // defaults of main.cu:393-400
    float alpha = 0.1f;
    bool write = false;
    unsigned long long seed = 1234ULL;
    static struct option long_options[] = {{"lattice-nbx", required_argument, 0, 'x'},
                                           {"lattice-nby", required_argument, 0, 'y'},
                                           {"ngpus", required_argument, 0, 'g'},
                                           {"seed", required_argument, 0, 's'},
                                           {"nwarmup", required_argument, 0, 'w'},
                                           {"niter", required_argument, 0, 'n'},
                                           {"write-lattice", no_argument, 0, 'o'},
                                           {"help", no_argument, 0, 'h'},
                                           {0, 0, 0, 0}};
    for (;;) {
        int option_index = 0;
        const int ch = getopt_long(argc, argv, "x:y:g:a:s:w:n:oh", long_options, &option_index);
        if (ch == -1) break;
        switch (ch) {
            case 0: break;
            case 'x': nbx = atoi(optarg); break;
            case 'y': nby = atoi(optarg); break;
            case 'g': nGPUs = atoi(optarg); break;
            case 'a': alpha = atof(optarg); break;
            case 's': seed = atoll(optarg); break;
            case 'w': nwarmup = atoi(optarg); break;
            case 'n': niter = atoi(optarg); break;
            case 'o': write = true; break;
            case 'h': usage(argv[0]); break;
            case '?': exit(EXIT_FAILURE);
            default: fprintf(stderr, "unknown option: %c\n", ch); exit(EXIT_FAILURE);
        }
    }
    if (nby % nGPUs != 0) {
        fprintf(stderr, "ERROR: Number of super blocks in y dimension must be multiple of number of gpus.\n");
        exit(EXIT_FAILURE);
    }
    const long long nx = (long long)nbx * 256, ny = (long long)nby * 256;

    ising_tc_ctx* ctx = nullptr;
    TC_CALL(ising_tc_create_multi(nbx, nby, seed, alpha, 0, nGPUs, &ctx));
    TC_CALL(ising_tc_init(ctx));
    TC_CALL(ising_tc_sync(ctx));

    printf("Starting warmup...\n");
    TC_CALL(ising_tc_update(ctx, 1, nwarmup));          // iterations 1..nwarmup (main.cu:582-585)
    TC_CALL(ising_tc_sync(ctx));
    printf("Starting trial iterations...\n");
    const auto t0 = std::chrono::high_resolution_clock::now();
    for (int n = nwarmup; n < niter + nwarmup; n++) {
        TC_CALL(ising_tc_update(ctx, (uint64_t)n + 1, 1));
        if ((n - nwarmup) % 1000 == 0) printf("Completed %d/%d iterations...\n", n - nwarmup + 1, niter);
    }
    TC_CALL(ising_tc_sync(ctx));
    const auto t1 = std::chrono::high_resolution_clock::now();
    const double duration = (double)std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count();

    printf("REPORT:\n");
    printf("\tnGPUs: %d\n", nGPUs);
    printf("\ttemperature: %f * %f\n", alpha, ISING_CRIT_TEMP);
    printf("\tseed: %llu\n", seed);
    printf("\twarmup iterations: %d\n", nwarmup);
    printf("\ttrial iterations: %d\n", niter);
    printf("\tlattice dimensions: %lld x %lld\n", nx, ny);
    printf("\telapsed time: %f sec\n", duration * 1e-6);
    printf("\tupdates per ns: %f\n", (double)(nx * ny) * niter / duration * 1e-3);

    int64_t sum = 0;
    TC_CALL(ising_tc_sum(ctx, &sum));
    std::cout << "\taverage magnetism (absolute): " << fabs((double)sum / (double)(nx * ny)) << std::endl;

    if (write) {
        printf("Writing lattice to %s...\n", "final.txt");
        TC_CALL(ising_tc_write_lattice(ctx, "final.txt"));
    }
    TC_CALL(ising_tc_destroy(ctx));
    return 0;
}
