/* Host-side partitioner call of the Cluster-GCN path: the same METIS invocation
 * torch_sparse.partition makes for PyG's ClusterData (/root/reference/data/dataset_handler.py:273):
 * METIS_PartGraphKway, ncon = 1, no vertex/edge weights, default options, 64-bit idx_t.
 * METIS is third-party library code (the static library ships in the CUDA toolkit); no metis.h is
 * installed, so the prototype is declared here. */
#include <stdint.h>
#include "lgcn_b200.h"

typedef int64_t idx_t;
typedef float real_t;
int METIS_PartGraphKway(idx_t *nvtxs, idx_t *ncon, idx_t *xadj, idx_t *adjncy, idx_t *vwgt, idx_t *vsize,
                        idx_t *adjwgt, idx_t *nparts, real_t *tpwgts, real_t *ubvec, idx_t *options,
                        idx_t *objval, idx_t *part);

int lgcn_partition_metis(int64_t num_nodes, const int64_t *indptr, const int64_t *index, int64_t num_parts,
                         int64_t *part_out) {
    if (!indptr || !index || !part_out || num_nodes <= 0 || num_parts <= 0) return LGCN_E_INVALID;
    if (num_parts == 1) {
        for (int64_t i = 0; i < num_nodes; ++i) part_out[i] = 0;
        return LGCN_OK;
    }
    idx_t nvtxs = num_nodes, ncon = 1, np = num_parts, objval = -1;
    int rc = METIS_PartGraphKway(&nvtxs, &ncon, (idx_t *)indptr, (idx_t *)index, 0, 0, 0, &np, 0, 0, 0, &objval,
                                 part_out);
    return rc == 1 ? LGCN_OK : LGCN_E_METIS;
}
