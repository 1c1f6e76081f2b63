/* ORACLE (test infrastructure).  Thin shim over the METIS 5.x static library that ships in
 * the CUDA toolkit (libmetis_static.a, 64-bit idx_t), making the same call as
 * torch_sparse/csrc/cpu/metis_cpu.cpp::partition_cpu (recursive=false, no weights):
 *   METIS_PartGraphKway(&nvtxs,&ncon=1,xadj,adjncy,NULL,NULL,NULL,&nparts,NULL,NULL,NULL,&objval,part)
 * reached from the reference at data/dataset_handler.py:273 (ClusterData).  No metis.h is
 * installed, so the prototype is declared by hand. */
#include <stdint.h>
typedef int64_t idx_t;
typedef float real_t;
int METIS_PartGraphKway(idx_t *nvtxs, idx_t *ncon, idx_t *xadj, idx_t *adjncy, idx_t *vwgt,
                        idx_t *vsize, idx_t *adjwgt, idx_t *nparts, real_t *tpwgts,
                        real_t *ubvec, idx_t *options, idx_t *objval, idx_t *part);

int oracle_metis_kway(int64_t n, int64_t *xadj, int64_t *adjncy, int64_t nparts, int64_t *part) {
    idx_t nvtxs = n, ncon = 1, np = nparts, objval = -1;
    return METIS_PartGraphKway(&nvtxs, &ncon, xadj, adjncy, 0, 0, 0, &np, 0, 0, 0, &objval, part);
}
