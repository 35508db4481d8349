// C ABI, part 2: simulation, null distribution, pair statistics, clustering.
#include "../../include/comap_b200.h"
#include "context.h"

namespace cmb { extern thread_local std::string g_last_error; }
using namespace cmb;
struct cmb_ctx { Context c; };

#define CMB_TRY try {
#define CMB_CATCH                                   \
  }                                                 \
  catch (const std::exception& e) {                 \
    g_last_error = e.what();                        \
    return 1;                                       \
  }                                                 \
  catch (...) {                                     \
    g_last_error = "unknown error";                 \
    return 1;                                       \
  }                                                 \
  return 0;
#define CMB_TODO(name) CMB_TRY fail(name ": not implemented yet"); CMB_CATCH

extern "C" {

int cmb_simulate(cmb_ctx*, uint64_t, int64_t, int64_t, int32_t, uint8_t*, int32_t*) { CMB_TODO("cmb_simulate") }
int cmb_null_intra(cmb_ctx*, int32_t, uint64_t, int32_t, int32_t, int32_t, int32_t, int32_t, int32_t, double, double*) { CMB_TODO("cmb_null_intra") }
int cmb_null_intra_from_alignments(cmb_ctx*, int32_t, int32_t, int32_t, const uint8_t*, const uint8_t*, int32_t, double, double*) { CMB_TODO("cmb_null_intra_from_alignments") }
int cmb_null_samples_dev(cmb_ctx*, const double**, const double**, int64_t*) { CMB_TODO("cmb_null_samples_dev") }
int cmb_null_load_dev(cmb_ctx*, const double*, const double*, int64_t, int32_t, double) { CMB_TODO("cmb_null_load_dev") }
int cmb_null_get(cmb_ctx*, int32_t*, double*, int64_t*, double*, int64_t) { CMB_TODO("cmb_null_get") }
int cmb_pairs(cmb_ctx*, int32_t, const cmb_filters*, int32_t, int32_t, int32_t, int64_t, int32_t*, int32_t*, double*, int32_t*, double*, double*, double*, int64_t*, int64_t*) { CMB_TODO("cmb_pairs") }
int cmb_distance_matrix(cmb_ctx*, int32_t, double*) { CMB_TODO("cmb_distance_matrix") }
int cmb_cluster(cmb_ctx*, int32_t, int32_t*, int32_t*, double*) { CMB_TODO("cmb_cluster") }
int cmb_groups(cmb_ctx*, int32_t, int32_t, int32_t*, int64_t*, double*, double*, double*, int64_t*) { CMB_TODO("cmb_groups") }
int cmb_cluster_null(cmb_ctx*, int32_t, int32_t, uint64_t, int32_t, int32_t, int32_t, int32_t, int64_t, int64_t, int32_t*, int32_t*, double*, double*, double*, int32_t*, int64_t*, int64_t*) { CMB_TODO("cmb_cluster_null") }

} // extern "C"
