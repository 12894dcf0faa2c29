// cuising_main.cpp -- command-line host of the B200 engine with the reference's CLI surface
// (optimized/main.cu:1285-1479 options and defaults, :1483-1594 banner, :1747/:1809/:1831/
// :1874/:1884 report lines, :1140-1209 lattice dump).  Pure C++ over the C ABI in
// include/ising_b200.h: no CUDA calls here.  Error policy of the reference (cudamacro.h:25-39):
// message on stderr, exit(EXIT_FAILURE).
#include <getopt.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <chrono>
#include <string>
#include <vector>

#include "ising_b200.h"

namespace {

constexpr int SPIN_X_WORD = 16;              // the reference's unit for its banner (main.cu:1243)
constexpr int X_MULT = 2048;                 // 2*SPIN_X_WORD*2*BLOCK_X*BMULT_X (main.cu:821)
constexpr int Y_MULT = 16;                   // BLOCK_Y*BMULT_Y
constexpr float ALPHA_DEF = 0.1f;            // main.cu:43
constexpr float MIN_TEMP = 0.05f * ISING_CRIT_TEMP;   // main.cu:44
constexpr double TGT_MAGN_MAX_DIFF = 1.0E-3; // main.cu:65
constexpr int MAX_EXP_TIME = 200;            // main.cu:67
constexpr int MIN_EXP_TIME = 152;            // main.cu:68
constexpr int MAX_CORR_LEN = 128;            // main.cu:70

#define ISING_CALL(call)                                                       \
    do {                                                                       \
        if ((call) != ISING_OK) {                                              \
            fprintf(stderr, "%s:%d: %s\n", __FILE__, __LINE__, ising_last_error()); \
            exit(EXIT_FAILURE);                                                \
        }                                                                      \
    } while (0)

[[noreturn]] void usage(const char* pname)
{
    const char* bname = strrchr(pname, '/');
    bname = bname ? bname + 1 : pname;
    printf("Usage: %s [options]\n"
           "options:\n"
           "\t-x|--x <HORIZ_DIM>\n"
           "\t\tSpecifies the horizontal dimension of the entire  lattice  (black+white  spins),\n"
           "\t\tper GPU. This dimension must be a multiple of %d.\n\n"
           "\t-y|--y <VERT_DIM>\n"
           "\t\tSpecifies the vertical dimension of the entire lattice (black+white spins),  per\n"
           "\t\tGPU. This dimension must be a multiple of %d.\n\n"
           "\t-n|--n <NSTEPS>\n"
           "\t\tSpecifies the number of iteration to run.\n"
           "\t\tDefualt: %d\n\n"
           "\t-d|--devs <NUM_DEVICES>\n"
           "\t\tSpecifies the number of GPUs to use. Will use devices with ids [0, NUM_DEVS-1].\n"
           "\t\tDefualt: 1.\n\n"
           "\t-s|--seed <SEED>\n"
           "\t\tSpecifies the seed used to generate random numbers.\n"
           "\t\tDefault: %llu\n\n"
           "\t-a|--alpha <ALPHA>\n"
           "\t\tSpecifies the temperature in T_CRIT units.  If both this  option  and  '-t'  are\n"
           "\t\tspecified then the '-t' option is used.\n"
           "\t\tDefault: %f\n\n"
           "\t-t|--temp <TEMP>\n"
           "\t\tSpecifies the temperature in absolute units.  If both this option and  '-a'  are\n"
           "\t\tspecified then this option is used.\n"
           "\t\tDefault: %f\n\n"
           "\t-p|--print <STAT_FREQ>\n"
           "\t\tSpecifies the frequency, in no.  of  iteration,  with  which  the  magnetization\n"
           "\t\tstatistics is printed.  If this option is used together to the '-e' option, this\n"
           "\t\toption is ignored.\n"
           "\t\tDefault: only at the beginning and at end of the simulation\n\n"
           "\t-e|--exppr\n"
           "\t\tPrints the magnetization at time steps in the series 0 <= 2^(x/4) < NSTEPS.   If\n"
           "\t\tthis option is used  together  to  the  '-p'  option,  the  latter  is  ignored.\n"
           "\t\tDefault: disabled\n\n"
           "\t-c|--corr\n"
           "\t\tDumps to a  file  named  corr_{X}x{Y}_T_{TEMP}  the  correlation  of each  point\n"
           "\t\twith the  %d points on the right and below.  The correlation is computed  every\n"
           "\t\ttime the magnetization is printed on screen (based on either the  '-p'  or  '-e'\n"
           "\t\toption) and it is written in the file one line per measure.\n"
           "\t\tDefault: disabled\n\n"
           "\t-m|--magn <TGT_MAGN>\n"
           "\t\tSpecifies the magnetization value at which the simulation is  interrupted.   The\n"
           "\t\tmagnetization of the system is checked against TGT_MAGN every STAT_FREQ, if  the\n"
           "\t\t'-p' option is specified, or according to the exponential  timestep  series,  if\n"
           "\t\tthe '-e' option is specified.  If neither '-p' not '-e' are specified then  this\n"
           "\t\toption is ignored.\n"
           "\t\tDefault: unset\n\n"
           "\t-J|--J <PROB>\n"
           "\t\tSpecifies the probability [0.0-1.0] that links  connecting  any  two  spins  are\n"
           "\t\tanti-ferromagnetic. \n"
           "\t\tDefault: 0.0\n\n"
           "\t   --xsl <HORIZ_SUB_DIM>\n"
           "\t\tSpecifies the horizontal dimension of each sub-lattice (black+white spins),  per\n"
           "\t\tGPU.  This dimension must be a divisor of the horizontal dimension of the entire\n"
           "\t\tlattice per  GPU  (specified  with  the  '-x'  option) and a multiple of %d.\n"
           "\t\tDefault: sub-lattices are disabled.\n\n"
           "\t   --ysl <VERT_SUB_DIM>\n"
           "\t\tSpecifies the vertical  dimension of each  sub-lattice (black+white spins),  per\n"
           "\t\tGPU.  This dimension must be a divisor of the vertical dimension of  the  entire\n"
           "\t\tlattice per  GPU  (specified  with  the  '-y'  option) and a multiple of %d.\n\n"
           "\t-o|--o\n"
           "\t\tEnables the file dump of  the lattice  every time  the magnetization is printed.\n"
           "\t\tDefault: off\n\n"
           "Options not in the reference:\n"
           "\t   --save <FILE>\n"
           "\t\tWrites a binary checkpoint (bit-packed lattice of all GPUs, seed,  iterations)\n"
           "\t\tat the end of the run.\n\n"
           "\t   --load <FILE>\n"
           "\t\tStarts from a checkpoint instead of the random lattice; iteration numbers -- and\n"
           "\t\ttherefore the random numbers -- continue where the saved run stopped.\n\n",
           bname, X_MULT, Y_MULT, 1, (unsigned long long)ISING_DEFAULT_SEED, ALPHA_DEF,
           ALPHA_DEF * ISING_CRIT_TEMP, MAX_CORR_LEN, X_MULT, Y_MULT);
    exit(EXIT_SUCCESS);   // the reference's usage() always exits with success (main.cu:828)
}

// Print times of the -e series (generate_times, main.cu:1211-1228): the first entry is
// MIN_EXP_TIME, then every t = rint(2^(j/4)) that at least doubles the previous entry.
std::vector<unsigned long long> exp_series(unsigned long long nsteps)
{
    std::vector<unsigned long long> t(MAX_EXP_TIME, 0ull);
    int nt = 0;
    t[0] = MIN_EXP_TIME;
    unsigned long long cur = 0;
    for (unsigned long long j = 0; j < nsteps && cur < nsteps; j++) {
        cur = (unsigned long long)rint(pow(2.0, j / 4.0));
        if (cur >= 2 * t[nt] && nt < MAX_EXP_TIME - 1) t[++nt] = cur;
    }
    return t;
}

}  // namespace

int main(int argc, char** argv)
{
    int X = 0, Y = 0, dumpOut = 0, corrOut = 0, nsteps = 1, ndev = 1;
    const char *ckptSave = nullptr, *ckptLoad = nullptr;
    unsigned long long seed = ISING_DEFAULT_SEED;
    float alpha = -1.0f, temp = -1.0f, tempUpdStep = 0;
    int tempUpdFreq = 0, printFreq = 0, printExp = 0, printExpCur = 0;
    double tgtMagn = -1.0;
    int useGenHamilt = 0;
    float hamiltPerc1 = 0.0f;
    int useSubLatt = 0, XSL = 0, YSL = 0, NSLX = 1, NSLY = 1;

    // same table as main.cu:1288-1307, including its quirks (--nit, --out, --help takes an argument)
    static struct option long_options[] = {
        {"x", required_argument, 0, 'x'},     {"y", required_argument, 0, 'y'},
        {"nit", required_argument, 0, 'n'},   {"seed", required_argument, 0, 's'},
        {"out", no_argument, 0, 'o'},         {"devs", required_argument, 0, 'd'},
        {"alpha", required_argument, 0, 'a'}, {"temp", required_argument, 0, 't'},
        {"print", required_argument, 0, 'p'}, {"update", required_argument, 0, 'u'},
        {"magn", required_argument, 0, 'm'},  {"exppr", no_argument, 0, 'e'},
        {"corr", no_argument, 0, 'c'},        {"J", required_argument, 0, 'J'},
        {"xsl", required_argument, 0, 1},     {"ysl", required_argument, 0, 2},
        {"save", required_argument, 0, 3},    {"load", required_argument, 0, 4},   // checkpoint (not in the reference)
        {"help", required_argument, 0, 'h'},  {0, 0, 0, 0}};

    for (;;) {
        int option_index = 0;
        const int och = getopt_long(argc, argv, "x:y:n:ohs:d:a:t:p:u:m:ecJ:r:", long_options, &option_index);
        if (och == -1) break;
        switch (och) {
            case 0: break;
            case 'x': X = atoi(optarg); break;
            case 'y': Y = atoi(optarg); break;
            case 'n': nsteps = atoi(optarg); break;
            case 'o': dumpOut = 1; break;
            case 'h': usage(argv[0]);
            case 's':
                seed = atoll(optarg);
                if (seed == 0) seed = ((getpid() * rand()) & 0x7FFFFFFFF);
                break;
            case 'd': ndev = atoi(optarg); break;
            case 'a': alpha = atof(optarg); break;
            case 't': temp = atof(optarg); break;
            case 'p': printFreq = atoi(optarg); break;
            case 'e': printExp = 1; break;
            case 'u': {   // -u STEP,FREQ
                char* a0 = strtok(optarg, ",");
                if (!a0) { fprintf(stderr, "cannot find temperature step in parameter...\n"); exit(EXIT_FAILURE); }
                char* a1 = strtok(NULL, ",");
                if (!a1) { fprintf(stderr, "cannot find iteration count in parameter...\n"); exit(EXIT_FAILURE); }
                tempUpdStep = atof(a0);
                tempUpdFreq = atoi(a1);
                printf("tempUpdStep: %f, tempUpdFreq: %d\n", tempUpdStep, tempUpdFreq);
                break;
            }
            case 'm': tgtMagn = atof(optarg); break;
            case 'c': corrOut = 1; break;
            case 'J':
                useGenHamilt = 1;
                hamiltPerc1 = atof(optarg);
                hamiltPerc1 = fminf(fmaxf(0.0f, hamiltPerc1), 1.0f);
                break;
            case 1: useSubLatt = 1; XSL = atoi(optarg); break;
            case 2: useSubLatt = 1; YSL = atoi(optarg); break;
            case 3: ckptSave = optarg; break;
            case 4: ckptLoad = optarg; break;
            case '?': exit(EXIT_FAILURE);
            default: fprintf(stderr, "unknown option: %c\n", och); exit(EXIT_FAILURE);
        }
    }

    // defaulting of missing dimensions, main.cu:1395-1410
    if (!X || !Y) {
        if (!X) X = (Y && !(Y % X_MULT)) ? Y : X_MULT;
        if (!Y) Y = !(X % Y_MULT) ? X : Y_MULT;
    }
    if (!X || (X % 2) || ((X / 2) % (X_MULT / 2))) {
        fprintf(stderr, "\nPlease specify an X dim multiple of %d\n\n", X_MULT);
        usage(argv[0]);
    }
    if (!Y || (Y % Y_MULT)) {
        fprintf(stderr, "\nPlease specify a Y dim multiple of %d\n\n", Y_MULT);
        usage(argv[0]);
    }
    if (useSubLatt) {   // main.cu:1423-1456
        if (!XSL || !YSL) {
            if (!XSL) XSL = (YSL && !(YSL % X_MULT)) ? YSL : X_MULT;
            if (!YSL) YSL = !(XSL % Y_MULT) ? XSL : Y_MULT;
        }
        if ((X % XSL) || !XSL || (XSL % 2) || ((XSL / 2) % (X_MULT / 2))) {
            fprintf(stderr, "\nPlease specify an X sub-lattice dim multiple of %d and divisor of %d\n\n", X_MULT, X);
            usage(argv[0]);
        }
        if ((Y % YSL) || !YSL || (YSL % Y_MULT)) {
            fprintf(stderr, "\nPlease specify a Y sub-lattice dim multiple of %d divisor of %d\n\n", Y_MULT, Y);
            usage(argv[0]);
        }
        NSLX = X / XSL;
        NSLY = Y / YSL;
    } else {
        XSL = X;
        YSL = Y * ndev;
    }
    if (temp == -1.0f) temp = (alpha == -1.0f) ? ALPHA_DEF * ISING_CRIT_TEMP : alpha * ISING_CRIT_TEMP;
    if (printExp && printFreq) printFreq = 0;
    std::vector<unsigned long long> printExpSteps;
    if (printExp) printExpSteps = exp_series((unsigned long long)nsteps);


    printf("\nUsing GPUs:\n");
    for (int i = 0; i < ndev; i++) {
        char name[256];
        int sms, thr, maj, min, ecc;
        ISING_CALL(ising_device_info(i, name, sizeof(name), &sms, &thr, &maj, &min, &ecc));
        printf("\t%2d (%s, %d SMs, %d th/SM max, CC %d.%d, ECC %s)\n", i, name, sms, thr, maj, min, ecc ? "on" : "off");
    }
    printf("\n");

    if (ndev > 1) {   // main.cu:1507-1536
        printf("GPUs direct access matrix:\n       ");
        for (int i = 0; i < ndev; i++) printf("%4d", i);
        int missingLinks = 0;
        printf("\n");
        for (int i = 0; i < ndev; i++) {
            printf("GPU %2d:", i);
            for (int j = 0; j < ndev; j++) {
                int access = 1;
                if (i != j) {
                    ISING_CALL(ising_can_access_peer(i, j, &access));
                    if (!access) missingLinks++;
                }
                printf("%4c", access ? 'V' : 'X');
            }
            printf("\n");
        }
        printf("\n");
        if (missingLinks) {
            fprintf(stderr, "error: %d direct memory links among devices missing\n", missingLinks);
            exit(EXIT_FAILURE);
        }
    }

    // the banner keeps the reference's units (16 spins per 64-bit word) so that scripts that
    // parse it keep working; the device packing of this engine is 1 bit per spin
    const size_t lld = (size_t)(X / 2) / SPIN_X_WORD;
    const size_t llenLoc = (size_t)Y * lld;
    const size_t llen = 2ull * ndev * llenLoc;
    const size_t nspins = llen * SPIN_X_WORD;

    printf("Run configuration:\n");
    printf("\tspin/word: %d\n", SPIN_X_WORD);
    printf("\tspins: %zu\n", nspins);
    printf("\tseed: %llu\n", seed);
    printf("\titerations: %d\n", nsteps);
    printf("\tblock (X, Y): %d, %d\n", 16, 16);
    printf("\ttile  (X, Y): %d, %d\n", 32, 16);
    printf("\tgrid  (X, Y): %d, %d\n", X / X_MULT, Y / Y_MULT);
    if (printFreq) printf("\tprint magn. every %d steps\n", printFreq);
    else if (printExp) printf("\tprint magn. following exponential series\n");
    else printf("\tprint magn. at 1st and last step\n");
    if ((printFreq || printExp) && tgtMagn != -1.0)
        printf("\tearly exit if magn. == %lf+-%lf\n", tgtMagn, TGT_MAGN_MAX_DIFF);
    printf("\ttemp: %f (%f*T_crit)\n", temp, temp / ISING_CRIT_TEMP);
    if (!tempUpdFreq) printf("\ttemp update not set\n");
    else printf("\ttemp update: %f / %d iterations\n", tempUpdStep, tempUpdFreq);
    if (useGenHamilt) printf("\tusing Hamiltonian buffer, setting links to -1 with prob %G\n", hamiltPerc1);
    else printf("\tnot using Hamiltonian buffer\n");
    printf("\n");
    if (useSubLatt) {
        printf("\tusing sub-lattices:\n");
        printf("\t\tno. of sub-lattices per GPU: %8d\n", NSLX * NSLY);
        printf("\t\tno. of sub-lattices (total): %8d\n", ndev * NSLX * NSLY);
        printf("\t\tsub-lattices size:           %7d x %7d\n\n", XSL, YSL);
    }
    printf("\tlocal lattice size:      %8d x %8d\n", Y, X);
    printf("\ttotal lattice size:      %8d x %8d\n", ndev * Y, X);
    printf("\tlocal lattice shape: 2 x %8d x %8zu (%12zu %s)\n", Y, lld, llenLoc * 2, "ulls");
    printf("\ttotal lattice shape: 2 x %8d x %8zu (%12zu %s)\n", ndev * Y, lld, llen, "ulls");
    // device memory actually used by this engine: 1 bit per spin
    printf("\tmemory: %.2lf MB (%.2lf MB per GPU)\n", (nspins / 8) / (1024.0 * 1024.0),
           (nspins / 8 / ndev) / (1024.0 * 1024.0));

    ising_cfg cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.X = X; cfg.Y = Y; cfg.ndev = ndev; cfg.first_device = 0; cfg.seed = seed; cfg.temp = temp;
    cfg.xsl = useSubLatt ? XSL : 0; cfg.ysl = useSubLatt ? YSL : 0;
    if (useGenHamilt) { cfg.flags |= ISING_F_BONDS; cfg.j_prob = hamiltPerc1; }
    ising_ctx* ctx = nullptr;
    ISING_CALL(ising_create(&cfg, &ctx));

    ISING_CALL(ising_init_lattice(ctx));
    uint64_t it0 = 0;                       // sweeps already done by the run a checkpoint comes from
    if (ckptLoad) {
        ISING_CALL(ising_load_checkpoint(ctx, ckptLoad, &it0));
        printf("\nLoaded checkpoint %s (%llu iterations done)\n", ckptLoad, (unsigned long long)it0);
    }
    uint64_t cntPos = 0, cntNeg = 0;
    ISING_CALL(ising_count_spins(ctx, &cntPos, &cntNeg));
    printf("\nInitial magnetization: %9.6lf, up_s: %12llu, dw_s: %12llu\n",
           fabs((double)cntPos - (double)cntNeg) / (double)nspins, (unsigned long long)cntPos,
           (unsigned long long)cntNeg);
    ISING_CALL(ising_sync(ctx));

    // -c: one line per measure appended to corr_{Y}x{X}_T_{TEMP}_{SEED} (main.cu:1661, :1129-1135);
    // the file name keeps the temperature of the start of the run, as in the reference
    char cname[256];
    snprintf(cname, sizeof(cname), "corr_%dx%d_T_%f_%llu", Y, X, temp, seed);
    if (corrOut) remove(cname);
    auto corr = [&](int it) {
        double c[MAX_CORR_LEN];
        ISING_CALL(ising_correlation(ctx, MAX_CORR_LEN, c));
        FILE* fp = fopen(cname, "a");
        if (!fp) { fprintf(stderr, "cannot open %s\n", cname); exit(EXIT_FAILURE); }
        fprintf(fp, "%10d", it);
        for (int i = 0; i < MAX_CORR_LEN; i++) fprintf(fp, " % -12G", c[i]);
        fprintf(fp, "\n");
        fclose(fp);
    };
    auto dump = [&](int it) {
        char fname[256];
        snprintf(fname, sizeof(fname), "lattice_%dx%d_T_%f_IT_%08d_", Y, X, temp, it);   // main.cu:1816
        ISING_CALL(ising_dump_lattice(ctx, fname));
    };

    // timed region = the whole step loop, statistics included, as in the reference
    // (main.cu:1756-1871); the kernels are asynchronous, only the count/dump points synchronise
    const auto t0 = std::chrono::steady_clock::now();
    int j;
    for (j = 0; j < nsteps; j++) {
        ISING_CALL(ising_sweeps(ctx, (int)it0 + j + 1, 1));
        const bool atFreq = printFreq && ((j + 1) % printFreq) == 0;
        const bool atExp = printExp && printExpSteps[printExpCur] == (unsigned long long)j;
        bool stop = false;
        if (atFreq) {
            ISING_CALL(ising_count_spins(ctx, &cntPos, &cntNeg));
            const double magn = fabs((double)cntPos - (double)cntNeg) / (double)nspins;
            printf("        magnetization: %9.6lf, up_s: %12llu, dw_s: %12llu (iter: %8d)\n", magn,
                   (unsigned long long)cntPos, (unsigned long long)cntNeg, j + 1);
            if (corrOut) corr(j + 1);
            if (dumpOut) dump(j + 1);
            if (tgtMagn != -1.0 && fabs(magn - tgtMagn) < TGT_MAGN_MAX_DIFF) stop = true;
        }
        if (!stop && atExp) {
            printExpCur++;
            ISING_CALL(ising_count_spins(ctx, &cntPos, &cntNeg));
            const double magn = fabs((double)cntPos - (double)cntNeg) / (double)nspins;
            printf("        magnetization: %9.6lf (^2: %9.6lf), up_s: %12llu, dw_s: %12llu (iter: %8d)\n", magn,
                   magn * magn, (unsigned long long)cntPos, (unsigned long long)cntNeg, j + 1);
            if (corrOut) corr(j + 1);
            if (dumpOut) dump(j + 1);
            if (tgtMagn != -1.0 && fabs(magn - tgtMagn) < TGT_MAGN_MAX_DIFF) stop = true;
        }
        if (stop) { j++; break; }
        if (tempUpdFreq && ((j + 1) % tempUpdFreq) == 0) {   // main.cu:1848-1860
            temp = fmaxf(MIN_TEMP, temp + tempUpdStep);
            printf("Changing temperature to %f\n", temp);
            float tab[5];
            ising_accept_table(temp, tab);
            for (int i = 0; i < 2; i++)
                for (int k = 0; k < 5; k++) printf("exp[%2d][%d]: %E\n", i ? 1 : -1, k, i ? tab[k] : tab[4 - k]);
            ISING_CALL(ising_set_temperature(ctx, temp));
        }
    }
    ISING_CALL(ising_sync(ctx));
    const double et = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();

    ISING_CALL(ising_count_spins(ctx, &cntPos, &cntNeg));
    printf("Final   magnetization: %9.6lf, up_s: %12llu, dw_s: %12llu (iter: %8d)\n\n",
           fabs((double)cntPos - (double)cntNeg) / (double)nspins, (unsigned long long)cntPos,
           (unsigned long long)cntNeg, j);

    // flips/ns as main.cu:1885.  BW: bytes this engine moves per colour step = source read +
    // own-colour read + own-colour write at 1 bit/spin (the reference prints the same three
    // terms for its 4 bit/spin layout, main.cu:1887-1890).
    const double bytes_per_color = 3.0 * (double)(nspins / 2) / 8.0;
    printf("Kernel execution time for %d update steps: %E ms, %.2lf flips/ns (BW: %.2lf GB/s)\n", j, et,
           (double)nspins * j / (et * 1.0E+6), (2.0 * j * bytes_per_color / 1.0E+9) / (et / 1.0E+3));

    if (dumpOut) dump(j);
    if (ckptSave) {
        ISING_CALL(ising_save_checkpoint(ctx, ckptSave, it0 + (uint64_t)j));
        printf("Saved checkpoint %s (%llu iterations done)\n", ckptSave, (unsigned long long)(it0 + j));
    }
    ISING_CALL(ising_destroy(ctx));
    return 0;
}
