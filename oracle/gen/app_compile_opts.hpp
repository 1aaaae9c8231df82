// What the reference's ./configure would write into its (read-only) src/
// directory (reference: configure:3-12); only the `binary` subprogram is enabled.
#define MARA_COMPILE_SUBPROGRAM_BOILERPLATE 0
#define MARA_COMPILE_SUBPROGRAM_PARTDOM     0
#define MARA_COMPILE_SUBPROGRAM_CLOUD       0
#define MARA_COMPILE_SUBPROGRAM_SEDOV       0
#define MARA_COMPILE_SUBPROGRAM_BINARY      1
#define MARA_COMPILE_SUBPROGRAM_AMRSAND     0
#define MARA_COMPILE_SUBPROGRAM_TEST        0
#define MARA_PREFERRED_THREAD_COUNT         8
