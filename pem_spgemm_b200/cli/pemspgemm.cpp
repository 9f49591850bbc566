// pemspgemm — drop-in command line of the reference program, on top of the C ABI.
//
//   pemspgemm <path/to/file.mtx> <0|1> [anything]
//
// mirrors /root/reference/spgemm.cu::main (:720-1568): argument handling (:722-725, :782-792,
// :1485), the stdout report (:794-806, :1406-1422), one appended row of
// ./pemspgemm_benchmark_result.csv with the same 14 columns (:1424-1450, README.md:51-53), and
// the sorted COO dump into /tmp/SPGEMM_RESULT_{NNZ,ROWS,COLS,VALS}.txt (:1527-1560).
// Differences, all on the forgiving side: a missing second argument means "0" (the reference
// calls atoi(NULL) and crashes), every CUDA/IO error is reported and exits 2 (the reference
// checks nothing), the file is parsed once, and A is converted once when B == A.
//
// Environment (the positional surface is unchanged): PEM_REPEAT (default 10, Makefile:34),
// PEM_WARMUP (default 1, spgemm.cu:712-714), PEM_DEVICE (default 0), PEM_DUMP_DIR (default /tmp),
// PEM_CSV (default ./pemspgemm_benchmark_result.csv), PEM_KEEP_EMPTY=1 to carry the reference's empty C'
// tiles through all three steps (PEM_OPT_KEEP_EMPTY_TILES; "C tiles" is the reference's count either way: it
// comes from one untimed tile-level symbolic pass), PEM_PANELS=n to multiply in n sequential tile-row panels (a C that does not
// fit the GPU in tiled form is produced, dumped and freed panel by panel; timings are summed over the panels).
#include <charconv>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <string>
#include <vector>

#include "pemspgemm.h"

static int env_int(const char* name, int dflt)
{
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

#define DIE_IF(rc, ctx, what)                                                              \
    do {                                                                                   \
        if ((rc) != PEM_OK) {                                                              \
            std::cerr << "pemspgemm: " << what << " failed (" << (rc) << "): "            \
                      << ((ctx) ? pem_last_error(ctx) : "") << "\n";                      \
            return 2;                                                                      \
        }                                                                                  \
    } while (0)

int main(int argc, char* argv[])
{
    if (argc <= 1 || argc > 4) {
        std::cout << "Provide a matrix market file path. Exiting.\n";
        return 1;
    }
    const int REPEAT = std::max(1, env_int("PEM_REPEAT", 10));
    const int WARMUP = std::max(0, env_int("PEM_WARMUP", 1));
    const bool aat = (argc == 4);
    const bool dump = (argc >= 3) && atoi(argv[2]) != 0;

    auto conv_start = std::chrono::high_resolution_clock::now();  // before file I/O, as spgemm.cu:760
    int32_t rows = 0, cols = 0, *I = nullptr, *J = nullptr;
    int64_t nnz = 0;
    double* V = nullptr;
    int sym = 0;
    char err[256] = {0};
    int rc = pem_mtx_read(argv[1], &rows, &cols, &nnz, &I, &J, &V, &sym, err, sizeof err);
    if (rc != PEM_OK) {
        std::cerr << "pemspgemm: cannot read " << argv[1] << ": " << err << "\n";
        return 2;
    }
    if (rows != cols && !aat) {
        std::cout << "input is rectangular. Only AAt is possible. Exiting.\n";
        return 1;
    }
    const int32_t b_rows = aat ? cols : rows, b_cols = aat ? rows : cols;
    std::cout << "MATRIX A\n" << "filepath: " << argv[1] << "\n" << "Rows: " << rows << "\n"
              << "Cols: " << cols << "\n" << "Nnz: " << nnz << "\n";
    std::cout << "MATRIX B\n" << "filepath: " << argv[1] << "\n" << "Rows: " << b_rows << "\n"
              << "Cols: " << b_cols << "\n" << "Nnz: " << nnz << "\n";

    pem_ctx* ctx = nullptr;
    rc = pem_ctx_create(&ctx, env_int("PEM_DEVICE", 0));
    if (rc != PEM_OK) {
        std::cerr << "pemspgemm: no usable CUDA device (" << rc << "); this engine has no CPU fallback\n";
        return 2;
    }
    if (env_int("PEM_KEEP_EMPTY", 0)) pem_ctx_set_option(ctx, PEM_OPT_KEEP_EMPTY_TILES, 1);

    pem_tiled *A = nullptr, *B = nullptr;
    pem_times ta = {}, tb = {};
    rc = pem_convert_coo(ctx, rows, cols, nnz, I, J, V, 0, &A, &ta);
    DIE_IF(rc, ctx, "conversion of A");
    if (aat) {
        rc = pem_tiled_transpose(ctx, A, &B);          // B = A^T from A's tiles: no second parse / upload / sort
        DIE_IF(rc, ctx, "conversion of B");
        tb.convert_kernel_ms = ta.convert_kernel_ms;
    } else {
        B = A;  // A^2: one conversion serves both operands
    }
    pem_ctx_sync(ctx);
    auto conv_end = std::chrono::high_resolution_clock::now();
    const double conv_ms = std::chrono::duration<double, std::milli>(conv_end - conv_start).count();
    pem_free_host(I); pem_free_host(J); pem_free_host(V);

    uint64_t flop = 0;
    rc = pem_count_flop(ctx, A, B, &flop);
    DIE_IF(rc, ctx, "flop count");

    pem_tiled_info ia, ib;
    pem_tiled_info_get(A, &ia); pem_tiled_info_get(B, &ib);
    std::cout << "\nstep1 tile-level symbolic (bitmap / hash accumulators), B tile columns: " << ib.tile_cols << "\n";
    std::cout << "\nstep2 pemSpGEMM\n" << "\nstep3 pemSpGEMM\n\n\n";

    const int PANELS = std::max(1, env_int("PEM_PANELS", 1));
    std::vector<int32_t> bounds((size_t)PANELS + 1);
    rc = pem_partition_panels(ctx, A, B, PANELS, bounds.data());
    DIE_IF(rc, ctx, "panel split");
    // "C tiles" as the reference counts them (spgemm.cu:1420): every structurally reachable tile of C',
    // empty ones included.  One untimed tile-level symbolic pass; the timed products drop empty tiles.
    int64_t ref_c_tiles = 0;
    {
        pem_ctx_set_option(ctx, PEM_OPT_KEEP_EMPTY_TILES, 1);
        for (int pn = 0; pn < PANELS; ++pn) {
            pem_result* S = nullptr;
            rc = pem_step1_symbolic(ctx, A, B, bounds[(size_t)pn], bounds[(size_t)pn + 1], &S);
            DIE_IF(rc, ctx, "tile-level symbolic pass");
            pem_result_info si;
            pem_result_info_get(S, &si);
            ref_c_tiles += si.tiles;
            pem_result_free(ctx, S);
        }
        pem_ctx_set_option(ctx, PEM_OPT_KEEP_EMPTY_TILES, env_int("PEM_KEEP_EMPTY", 0) ? 1 : 0);
    }
    pem_result* C = nullptr;                      // the last panel of the last iteration (the whole C when PANELS == 1)
    int64_t c_tiles = 0, c_nnz = 0;
    double s1 = 0, s2 = 0, s3 = 0, total = 0, kernel = 0, mall = 0;
    for (int n = 0; n < WARMUP + REPEAT; ++n) {
        c_tiles = c_nnz = 0;
        for (int pn = 0; pn < PANELS; ++pn) {
            if (C) { pem_result_free(ctx, C); C = nullptr; }
            pem_times t = {};
            rc = pem_spgemm_panel(ctx, A, B, bounds[(size_t)pn], bounds[(size_t)pn + 1], &C, &t);
            DIE_IF(rc, ctx, "spgemm");
            pem_result_info pi;
            pem_result_info_get(C, &pi);
            c_tiles += pi.tiles; c_nnz += pi.nnz;
            if (n >= WARMUP) {
                s1 += t.step1_ms; s2 += t.step2_ms; s3 += t.step3_ms;
                total += t.total_ms; kernel += t.kernel_ms; mall += t.malloc_ms;
            }
        }
    }
    s1 /= REPEAT; s2 /= REPEAT; s3 /= REPEAT; total /= REPEAT; kernel /= REPEAT; mall /= REPEAT;
    std::cout << "warm up " << WARMUP << " time\n" << "average over " << REPEAT << " iterations\n\n";

    pem_result_info ic;
    pem_result_info_get(C, &ic);
    ic.tiles = c_tiles; ic.nnz = c_nnz;           // totals over the panels
    const double gflops = flop * 2.0 / (total * 1e6);                 // spgemm.cu:1403
    const double compression = ic.nnz ? (double)flop / (double)ic.nnz : 0.0;  // spgemm.cu:1404

    std::cout << std::fixed << std::setprecision(2);
    std::cout << "<---Program done--->\n";
    std::cout << "Matrix A CSR to tile kernel took---------" << ta.convert_kernel_ms << "ms\n";
    std::cout << "Matrix B CSR to tile kernel took---------" << tb.convert_kernel_ms << "ms\n";
    std::cout << "total conversion overhead----------------" << conv_ms << "ms\n\n";
    std::cout << "step1 - High Level Multiplication took---" << s1 << "ms\n";
    std::cout << "step2 - Allocating C took----------------" << s2 << "ms\n";
    std::cout << "step3 - Accumulation took----------------" << s3 << "ms\n\n";
    std::cout << "pemSpGEMM took " << total << "ms ----- GFlops: " << gflops << "\nKernel time " << kernel
              << "ms\nmalloc time " << mall << "ms\n";
    std::cout << "Flop count: " << flop << "\n\n";
    std::cout << "C tiles: " << ref_c_tiles << " (non-empty: " << ic.tiles << ")\n";
    std::cout << "C nnz: " << ic.nnz << "\n";
    std::cout << "Compression ratio " << compression << "\n";

    {   // CSV row: newline first, no header, std::fixed with 2 decimals (spgemm.cu:1424-1450)
        const char* csv_env = getenv("PEM_CSV");
        std::string csv = csv_env && *csv_env ? csv_env : "./pemspgemm_benchmark_result.csv";
        std::string path = argv[1], name = path;
        size_t slash = name.find_last_of('/');
        if (slash != std::string::npos) name = name.substr(slash + 1);
        size_t ext = name.rfind(".mtx");
        if (ext != std::string::npos) name = name.substr(0, ext);
        std::ofstream w(csv, std::ios::app);
        w << std::fixed << std::setprecision(2);
        w << "\n" << name << "," << flop << "," << ic.nnz << "," << compression << "," << ta.convert_kernel_ms << ","
          << tb.convert_kernel_ms << "," << conv_ms << "," << s1 << "," << s2 << "," << s3 << "," << total << ","
          << kernel << "," << mall << "," << gflops;
    }

    int exit_code = 0;
    if (!dump) {
        std::cout << "Not saving results. Exiting.\n";
    } else {
        const char* dd = getenv("PEM_DUMP_DIR");
        std::string dir = dd && *dd ? dd : "/tmp";
        std::cout << "Saving results to " << dir << "/SPGEMM_RESULT_*.txt\n";
        std::ofstream out;
        out.open(dir + "/SPGEMM_RESULT_NNZ.txt");
        out << ic.nnz;                                   // no trailing newline (spgemm.cu:1546)
        out.close();
        // one value per line, formatted and written by all host threads inside the library (pem_write_lines_*: the
        // reference streams one `<<` per line, spgemm.cu:1549-1558, which is slower than the SpGEMM by orders of magnitude)
        const std::string fr = dir + "/SPGEMM_RESULT_ROWS.txt", fc = dir + "/SPGEMM_RESULT_COLS.txt", fv = dir + "/SPGEMM_RESULT_VALS.txt";
        bool written = true;
        for (int pn = 0; pn < PANELS; ++pn) {            // panels in order = rows in order
            if (PANELS > 1) {                             // only the last panel is still alive: recompute the others
                pem_result_free(ctx, C); C = nullptr;
                rc = pem_spgemm_panel(ctx, A, B, bounds[(size_t)pn], bounds[(size_t)pn + 1], &C, nullptr);
                DIE_IF(rc, ctx, "spgemm");
            }
            pem_result_info pi;
            pem_result_info_get(C, &pi);
            std::vector<int32_t> r((size_t)pi.nnz), c((size_t)pi.nnz);
            std::vector<double> v((size_t)pi.nnz);
            rc = pem_result_to_coo(ctx, C, r.data(), c.data(), v.data());
            DIE_IF(rc, ctx, "COO export");
            written = pem_write_lines_i32(fr.c_str(), r.data(), pi.nnz, pn > 0) == PEM_OK && written;     // panels after the first append
            written = pem_write_lines_i32(fc.c_str(), c.data(), pi.nnz, pn > 0) == PEM_OK && written;
            written = pem_write_lines_f64(fv.c_str(), v.data(), pi.nnz, pn > 0) == PEM_OK && written;
        }
        if (!written) exit_code = 2;
    }
    std::cout << "CLEANING UP RESOURCES\n\n";
    pem_result_free(ctx, C);
    if (B != A) pem_tiled_free(ctx, B);
    pem_tiled_free(ctx, A);
    pem_ctx_destroy(ctx);
    return exit_code;
}
