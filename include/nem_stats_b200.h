/*
 * nem_stats_b200.h -- the reference's own native entry point, same symbol and signature, on the B200 path.
 *
 * /root/reference/ppanggolin/nem/NEM/nem_exe.h:23-38 declares
 *     extern int nem(const char* Fname, const int nk, const char* algo, const float beta, const char* convergence,
 *                    const float convergence_th, const char* format, const int it_max, const int dolog,
 *                    const char* model_family, const char* proportion, const char* dispersion, const int init_mode,
 *                    const char* init_file, const char* out_file_prefix, const int seed);
 * and NEM/nem_stats.pyx:1-17 binds exactly this as `nem_stats.nem`.  libnemb200.so exports the same function: it reads
 * the files the reference reads (<Fname>.str/.dat/.nei, init_file), runs the CUDA path (nemb200_nem_host) and writes the
 * files the reference writes (<out_file_prefix>.uf/.mf/.stderr) in the reference's formats (nem_exe.c:1601-1786), so an
 * UNMODIFIED ppanggolin/nem/partition.py works against it (a Cython shim identical to nem_stats.pyx, or the ctypes
 * binding in INTEGRATION.md).  It keeps the text round trip, so it is the compatibility / diffing entry, not the fast one.
 *
 * Supported configuration = what PPanGGOLiN passes (partition.py:87-150): algo "nem", convergence "clas",
 * format "fuzzy", model_family "bern", proportion "pk", dispersion "sk_" | "skd", init_mode 2 with the init file of
 * partition.py:65-85.  Anything else returns 2 (EXIT_E_ARGS) / 3 (EXIT_E_FILE) like the reference's argument and file
 * errors (NEM/lib_io.h:25-32).  Return 0 = results written, 1 = empty class (no .uf/.mf written, nem_exe.c:629-636).
 */
#ifndef NEM_STATS_B200_H
#define NEM_STATS_B200_H
#ifdef __cplusplus
extern "C" {
#endif
int nem(const char* Fname, const int nk, const char* algo, const float beta, const char* convergence,
        const float convergence_th, const char* format, const int it_max, const int dolog, const char* model_family,
        const char* proportion, const char* dispersion, const int init_mode, const char* init_file,
        const char* out_file_prefix, const int seed);
#ifdef __cplusplus
}
#endif
#endif
