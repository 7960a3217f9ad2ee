/*
 * qsv_b200.h -- C ABI of the B200-native state-vector engine (libqsv_b200.so).
 *
 * This is the drop-in boundary for the hot path of PennyLane's lightning.gpu
 * (reference tree: /root/reference, paths below relative to
 * pennylane_lightning_gpu/src/).  Every entry point replaces the body of a
 * reference method that today forwards to cuStateVec / cuSPARSE / cuBLAS /
 * CUDA-aware MPI; the reference-side binding is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain C: opaque handles, pointers and sizes only; no C++/torch types.
 *   - every function returns 0 on success, non-zero on failure;
 *     qsv_last_error() returns the message of the calling thread's last failure
 *     (the C++ shim rethrows it as Pennylane::Util::LightningException).
 *   - wires are PennyLane wires: wire w <-> amplitude-index bit (n-1-w)
 *     (simulator/StateVectorCudaManaged.hpp:1411-1418).
 *   - matrices are row-major in PennyLane wire order (first listed wire = most
 *     significant matrix bit), interleaved (re,im) DOUBLES for both precisions
 *     (they are a few hundred bytes; the kernels down-convert for complex64).
 *   - host array arguments are borrowed for the duration of the call only.
 *   - results are valid on return (stream-ordered inside, one sync at read-back).
 *   - thread-compatible: no shared mutable globals; one thread per handle at a time.
 */
#ifndef QSV_B200_H
#define QSV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qsv_state qsv_state; /* one 2^n amplitude array on one GPU      */
typedef struct qsv_ops qsv_ops;     /* a recorded circuit (OpsData equivalent)  */
typedef struct qsv_obs qsv_obs;     /* an observable tree (ObservableGPU<T>)    */

enum { QSV_C64 = 0, QSV_C128 = 1 }; /* cuFloatComplex / cuDoubleComplex         */

/* ---- library ------------------------------------------------------------- */
const char *qsv_last_error(void);
int qsv_version(void);
/* util/cuda_helpers.hpp:601-643 (getGPUCount / getGPUIdx / getGPUArch / is_gpu_supported) */
int qsv_device_count(int *count);
int qsv_device_arch(int device, int *major, int *minor);
int qsv_device_mem_info(int device, size_t *free_bytes, size_t *total_bytes);
/* deviceReset (cuda_helpers.hpp) and the allToAllAccess loop of bindings/Bindings.cpp:1737-1741 */
int qsv_device_reset(void);
int qsv_enable_peer_access(void);

/* ---- state-vector lifetime: StateVectorCudaManaged ctor/dtor (Managed.hpp:83-134),
 *      DataBuffer (util/DataBuffer.hpp:31-121) ------------------------------- */
int qsv_create(int n_qubits, int dtype, int device, qsv_state **out);
/* wrap caller-owned device memory (e.g. a torch tensor) and/or a caller stream */
int qsv_create_external(int n_qubits, int dtype, int device, void *device_ptr,
                        void *cuda_stream, qsv_state **out);
int qsv_destroy(qsv_state *sv);
int qsv_set_stream(qsv_state *sv, void *cuda_stream);
int qsv_synchronize(qsv_state *sv);
void *qsv_data_ptr(qsv_state *sv);
int qsv_num_qubits(const qsv_state *sv);
int qsv_dtype(const qsv_state *sv);
int qsv_device(const qsv_state *sv);

/* ---- initialisation and copies ------------------------------------------- */
/* initSV / setBasisState (StateVectorCudaBase.hpp:234-240, Managed.hpp:144-152, initSV.cu:107-117) */
int qsv_set_basis_state(qsv_state *sv, uint64_t index);
/* setStateVector (Managed.hpp:164-185, initSV.cu:64-97): zero, then sv[idx[i]] = val[i];
 * values are in the state's own precision (interleaved re,im) */
int qsv_set_state_vector(qsv_state *sv, const int64_t *indices, const void *values, size_t count);
/* CopyHostDataToGpu / CopyGpuDataToHost / CopyGpuDataToGpuIn (StateVectorCudaBase.hpp:104-203);
 * host data in the state's own precision, n_amps <= 2^n */
int qsv_h2d(qsv_state *sv, const void *host, size_t n_amps);
int qsv_d2h(qsv_state *sv, void *host, size_t n_amps);
int qsv_d2d(qsv_state *dst, const qsv_state *src);

/* ---- gates ----------------------------------------------------------------
 * applyOperation(name, wires, adjoint, params, matrix) (Managed.hpp:198-247) and the 35 named
 * apply* methods (Managed.hpp:321-560).  Unknown names with a matrix take the matrix path; unknown
 * names without one fail with "Currently unsupported gate: <name>" (Managed.hpp:238-239). */
int qsv_apply_named(qsv_state *sv, const char *name, const int *wires, int n_wires, int adjoint,
                    const double *params, int n_params);
/* applyDeviceMatrixGate / applyHostMatrixGate (Managed.hpp:1400-1568): 2^nt x 2^nt matrix on
 * targets, conditioned on all controls being 1 */
int qsv_apply_matrix(qsv_state *sv, const double *matrix_re_im, const int *ctrl_wires, int n_ctrls,
                     const int *tgt_wires, int n_tgts, int adjoint);
/* 10 applyGenerator* (Managed.hpp:563-688) + GateGenerators.hpp:58-321: state <- G state;
 * *scale receives the factor of AdjointDiffGPU.hpp:96-114 */
int qsv_apply_generator(qsv_state *sv, const char *name, const int *wires, int n_wires, int adjoint,
                        double *scale);

/* ---- recorded circuits: OpsData<SV> (AdjointDiffGPU.hpp:178-204, bindings/Bindings.cpp:806-840) */
int qsv_ops_create(qsv_ops **out);
int qsv_ops_destroy(qsv_ops *ops);
/* matrix_re_im may be NULL; mat_dim = 2^n_wires when given */
int qsv_ops_append(qsv_ops *ops, const char *name, const int *wires, int n_wires,
                   const double *params, int n_params, int inverse, const double *matrix_re_im,
                   size_t mat_dim);
int qsv_ops_size(const qsv_ops *ops);
/* whole-circuit application: consecutive gates are fused on the host and executed by the
 * shared-memory tile kernels (one HBM sweep per fused block); fuse = 0 applies gate by gate.
 * This is the batched form of apply_cq (lightning_gpu.py:519-555). */
int qsv_apply_ops(qsv_state *sv, const qsv_ops *ops, int fuse);
/* host-only view of what qsv_apply_ops(fuse = 1) would do for an n_qubits register (no device needed): gates are
 * merged, then packed into HBM sweeps over the dependency DAG of the circuit (dag = 1) or in program order
 * (dag = 0); low_bits = contiguous low tile bits (0 = default).  order_valid = 1 when every gate is executed
 * exactly once and every pair of gates that does not commute structurally keeps its program order. */
int qsv_ops_plan_sweeps(const qsv_ops *ops, int n_qubits, int dag, int low_bits, int64_t *n_gates_merged,
                        int64_t *n_sweeps, int64_t *max_gates_per_sweep, int *order_valid);
/* host-only: the arithmetic apply_ops(fuse = 1) performs for this circuit with the default planner settings -- fused
 * multiply-adds per amplitude (multiply by 2^n_qubits for the whole state), HBM sweeps and register passes */
int qsv_ops_plan_work(const qsv_ops *ops, int n_qubits, int dtype, double *fma_per_amplitude, int64_t *n_sweeps,
                      int64_t *n_passes);
/* statistics of the last qsv_apply_ops on this state: kernel launches and HBM sweeps */
int qsv_last_apply_stats(const qsv_state *sv, int64_t *launches, int64_t *sweeps);

/* ---- measurements: MeasurementsGPU part of Managed.hpp:702-1148 ------------ */
/* expval(name, wires, params, matrix) (Managed.hpp:702-751); out = {re, im} */
int qsv_expval_named(qsv_state *sv, const char *name, const int *wires, int n_wires,
                     const double *params, int n_params, double *out_re_im);
/* expval(wires, matrix) (Managed.hpp:755-780, :1577-1713) */
int qsv_expval_matrix(qsv_state *sv, const double *matrix_re_im, const int *wires, int n_wires,
                      double *out_re_im);
/* getExpectationValuePauliWords (Managed.hpp:1071-1148): n_terms words, word t is
 * letters[offsets[t]..offsets[t+1]) over {I,X,Y,Z} on wires[offsets[t]..offsets[t+1]);
 * per_term (may be NULL) receives <psi|P_t|psi> as double; *out = Re sum_t coeff_t <P_t> */
int qsv_expval_pauli_words(qsv_state *sv, int n_terms, const char *letters, const int *wires,
                           const int *offsets, const double *coeffs_re_im, double *per_term,
                           double *out);
/* getExpectationValueOnSparseSpMV (Managed.hpp:795-922): CSR with index_bytes 4 or 8,
 * values interleaved (re,im) doubles; *out = Re <psi|H psi> */
int qsv_expval_csr(qsv_state *sv, const void *row_offsets, const void *col_indices,
                   const double *values_re_im, int64_t nnz, int index_bytes, double *out);
/* probability(wires) (Managed.hpp:931-970): 2^k doubles, FIRST listed wire = LSB of the output
 * index (the reference's cuStateVec bit order; lightning_gpu.py:920-924 re-transposes) */
int qsv_probs(qsv_state *sv, const int *wires, int n_wires, double *out);
/* generate_samples (Managed.hpp:982-1061): inverse-CDF sampling from caller-supplied uniform
 * numbers in [0,1) (the reference draws them from mt19937 on the host, :1003-1007);
 * out[shot * n + w] = bit of wire w */
int qsv_sample(qsv_state *sv, const double *uniforms, int64_t shots, uint64_t *out);
/* innerProdC_CUDA (util/cuda_helpers.hpp:456-497): <a|b> */
int qsv_inner_product(qsv_state *a, qsv_state *b, double *out_re_im);
/* scaleAndAddC_CUDA (cuda_helpers.hpp:512-528): y += alpha * x */
int qsv_axpy(const double *alpha_re_im, const qsv_state *x, qsv_state *y);

/* ---- observables: ObservableGPU<T> family (algorithms/ObservablesGPU.hpp:56-587) */
int qsv_obs_named(const char *name, const int *wires, int n_wires, const double *params,
                  int n_params, qsv_obs **out);
int qsv_obs_hermitian(const double *matrix_re_im, size_t mat_dim, const int *wires, int n_wires,
                      qsv_obs **out);
int qsv_obs_tensor(qsv_obs *const *children, int n_children, qsv_obs **out); /* children are copied */
int qsv_obs_hamiltonian(const double *coeffs, qsv_obs *const *children, int n_children,
                        qsv_obs **out);
int qsv_obs_sparse(const int64_t *row_offsets, int64_t n_rows_plus_1, const int64_t *col_indices,
                   const double *values_re_im, int64_t nnz, qsv_obs **out);
int qsv_obs_destroy(qsv_obs *obs);
/* applyInPlace (ObservablesGPU.hpp:56): sv <- O sv */
int qsv_obs_apply(const qsv_obs *obs, qsv_state *sv);
/* Re <sv|O|sv> without modifying sv */
int qsv_obs_expval(const qsv_obs *obs, qsv_state *sv, double *out);

/* ---- adjoint Jacobian: AdjointJacobianGPU<T>::adjointJacobian (AdjointDiffGPU.hpp:499-596)
 * sv holds the final state (or the initial one when apply_operations != 0);
 * trainable indexes parametric ops only; jac is row-major [n_obs][n_trainable] doubles. */
int qsv_adjoint_jacobian(qsv_state *sv, const qsv_ops *ops, qsv_obs *const *observables, int n_obs,
                         const int64_t *trainable, int n_trainable, int apply_operations,
                         double *jac);

/* ---- sharded state vectors: StateVectorCudaMPI (simulator/StateVectorCudaMPI.hpp) with NCCL
 * send/recv in place of CUDA-aware MPI + custatevecSVSwapWorker (simulator/MPIWorker.hpp).
 * One process per GPU; the NCCL unique id is created on rank 0 and passed around by the host
 * language's own rendezvous (torch.distributed / MPI_Bcast). */
int qsv_dist_unique_id(void *id128);            /* 128 bytes out (rank 0)                    */
int qsv_dist_init(qsv_state *local, const void *id128, int rank, int world_size);
int qsv_dist_finalize(qsv_state *local);
/* global<->local index-bit swap (MPI.hpp:2488-2587): exchanges, with rank ^ (1 << (global_bit - n_local)),
 * the half of the local shard whose local_bit differs from this rank's value of global_bit (both are
 * PHYSICAL index bits), through chunk_bytes-sized staging buffers (0 = 256 MiB; the reference's
 * mpi_buf_size, MPIWorker.hpp:276-294), and records the exchange in the logical->physical qubit map */
int qsv_dist_swap_bits(qsv_state *local, int global_bit, int local_bit, size_t chunk_bytes);
/* whole circuit on the sharded register (wires refer to all n_total qubits): dense targets on global
 * qubits are swapped in lazily (farthest-next-use eviction), controls / diagonal gates on global qubits
 * run without communication (the reference swaps for every global wire, MPI.hpp:2023-2088) */
int qsv_dist_apply_ops(qsv_state *local, const qsv_ops *ops, int fuse, size_t chunk_bytes);
/* restore the identity qubit map so that rank r's shard is amplitudes [r * 2^n_local, (r+1) * 2^n_local) */
int qsv_dist_canonicalize(qsv_state *local, size_t chunk_bytes);
int qsv_dist_qubit_map(const qsv_state *local, int *phys_of_logical_bit, int n_total);
/* getExpectationValuePauliWords on a sharded register (MPI.hpp:1296-1445): local reduction + all-reduce */
int qsv_dist_expval_pauli_words(qsv_state *local, int n_terms, const char *letters, const int *wires,
                                const int *offsets, const double *coeffs_re_im, double *per_term, double *out);
/* setBasisState / setStateVector on the sharded register (MPI.hpp:332-381): indices address the whole register; the
 * qubit map is reset to the identity; every rank passes the same arguments and keeps the amplitudes of its shard */
int qsv_dist_set_basis_state(qsv_state *local, uint64_t index);
int qsv_dist_set_state_vector(qsv_state *local, const int64_t *indices, const void *values, size_t count);
/* expval(name, wires, ...) / expval(wires, matrix) on the sharded register (MPI.hpp:957-1035): target qubits are made
 * local, local reduction + all-reduce; out = {re, im}, identical on all ranks */
int qsv_dist_expval_named(qsv_state *local, const char *name, const int *wires, int n_wires, const double *params,
                          int n_params, double *out_re_im);
int qsv_dist_expval_matrix(qsv_state *local, const double *matrix_re_im, const int *wires, int n_wires, double *out_re_im);
/* getExpectationValueOnSparseSpMV on the sharded register (MPI.hpp:1050-1176, util/CSRMatrix.hpp:91-216): every rank
 * passes the whole CSR matrix (int64 indices) and uploads only its own row block; x is gathered from the other
 * shards over NVLink through their peer mappings (all-gather fallback without peer access) */
int qsv_dist_expval_csr(qsv_state *local, const int64_t *row_offsets, const int64_t *col_indices,
                        const double *values_re_im, int64_t nnz, double *out);
/* probability(wires) on the sharded register (MPI.hpp:1187-1290); same output convention as qsv_probs */
int qsv_dist_probs(qsv_state *local, const int *wires, int n_wires, double *out);
/* generate_samples on the sharded register (MPI.hpp:1454-1595): same definition as qsv_sample over the amplitude order
 * of the whole register; out[shot * n_total + w], identical on all ranks */
int qsv_dist_sample(qsv_state *local, const double *uniforms, int64_t shots, uint64_t *out);
/* ObservableGPUMPI<T> family (algorithms/ObservablesGPUMPI.hpp): wires of the observable address the whole register */
int qsv_dist_obs_expval(const qsv_obs *obs, qsv_state *local, double *out);
int qsv_dist_obs_apply(const qsv_obs *obs, qsv_state *local);
/* AdjointJacobianGPUMPI::adjointJacobian (algorithms/AdjointDiffGPUMPI.hpp:248-437): lambda and the bras are sharded
 * like the register and follow every exchange; one all-reduce of the Jacobian at the end; jac identical on all ranks */
int qsv_dist_adjoint_jacobian(qsv_state *local, const qsv_ops *ops, qsv_obs *const *observables, int n_obs,
                              const int64_t *trainable, int n_trainable, int apply_operations, double *jac);
int qsv_dist_rank(const qsv_state *local);
int qsv_dist_world_size(const qsv_state *local);
int qsv_dist_total_qubits(const qsv_state *local);
/* CopyHostDataToGpu / CopyGpuDataToHost / updateData of StateVectorCudaMPI (StateVectorCudaBase.hpp:104-228): the local
 * shard in the canonical layout (h2d resets the qubit map, d2h canonicalises first); copy = shard + qubit map */
int qsv_dist_h2d(qsv_state *local, const void *host, size_t n_amps);
int qsv_dist_d2h(qsv_state *local, void *host, size_t n_amps);
int qsv_dist_copy(qsv_state *dst, const qsv_state *src);
/* MPIManager::Barrier / Bcast / Scatter (util/MPIManager.hpp) over the register's NCCL communicator */
int qsv_dist_barrier(qsv_state *local);
int qsv_dist_bcast_bytes(qsv_state *local, void *host, size_t bytes, int root);
int qsv_dist_scatter_host(qsv_state *local, const void *send_host, void *recv_host, size_t bytes_per_rank, int root);
int qsv_dist_nccl_version(int *version);
/* MPI_Allreduce(sum) of small host vectors (MPI.hpp:1176, :1426, :2361): ncclAllReduce on a device buffer */
int qsv_dist_allreduce_f64(qsv_state *local, double *host_values, int count);
/* NVLink bytes sent by this rank and device milliseconds of the last exchange / of all exchanges */
int qsv_dist_last_swap_stats(const qsv_state *local, uint64_t *bytes_sent, float *ms);
int qsv_dist_total_swap_stats(qsv_state *local, int *n_swaps, uint64_t *bytes_sent, float *ms, int reset);
/* 1 when the exchanges run as direct NVLink load/store kernels on CUDA-IPC peer mappings (the default; NCCL
 * then only carries the handshakes), 0 when they fall back to staged NCCL send/recv (QSV_DIST_P2P=0) */
int qsv_dist_uses_peer_access(const qsv_state *local);
/* exchanges done through the second buffer (default where it fits, QSV_DIST_FUSED_SWAP=0 switches it off: the sweep before an
 * exchange stores out of place,
 * half of its tiles straight into the partner's buffer over NVLink) and how many of them a gate sweep carried; they are
 * not part of the swap statistics above, which time the in-place exchanges */
int qsv_dist_fused_exchange_stats(const qsv_state *local, int *n_out_of_place, int *n_carried_by_sweeps);
/* exchanges that were SPLIT between the sweep before (push a quarter of the shard, park a quarter) and the first sweep of the
 * next batch (fetch what the partner parked; QSV_DIST_SPLIT_XCHG): second halves performed, and how many a sweep carried */
int qsv_dist_split_exchange_stats(const qsv_state *local, int *n_second_halves, int *n_carried_by_sweeps);
/* Qubit-map policy of a sharded register.  lazy = 0 (default; the reference's contract, StateVectorCudaMPI keeps every
 * gate's swaps paired, MPI.hpp:2533-2583): every collective entry point returns with the canonical layout, so
 * qsv_dist_d2h and the Python `state` property are purely LOCAL copies of the shard (StateVectorCudaBase.hpp:104-228)
 * and a rank-conditional read cannot dead-lock.  lazy = 1: the logical->physical map persists between calls (a
 * swapped-in qubit stays local until evicted; what bench.py times); qsv_dist_d2h and qsv_dist_canonicalize are then
 * COLLECTIVE -- every rank must call them.  Collective itself (it restores the canonical layout when switching back). */
int qsv_dist_set_lazy_map(qsv_state *sv, int lazy);

/* host-only (no GPU, no NCCL): the exchanges qsv_dist_apply_ops would perform.  steps receives triples
 * (kind, a, b): kind 0 = swap physical global bit a with local bit b, kind 1 = apply op number a */
int qsv_dist_plan(const qsv_ops *ops, int n_total, int n_local, int *steps, int max_steps, int *n_steps,
                  int *final_phys_of_logical_bit);
/* the same, starting from a given qubit map (initial_phys_of_logical_bit[n_total], NULL = identity): what a second
 * qsv_dist_apply_ops on the same register would do, since the map persists between calls */
int qsv_dist_plan_from(const qsv_ops *ops, int n_total, int n_local, const int *initial_phys_of_logical_bit, int *steps,
                       int max_steps, int *n_steps, int *final_phys_of_logical_bit);

#ifdef __cplusplus
}
#endif
#endif /* QSV_B200_H */
