// Internal: communicator used by the Krylov driver (NCCL resolved at run time with dlopen so that
// the library loads - and the single-GPU path runs - without NCCL on the link line).
#pragma once
#include "tefem.h"
#include <cuda_runtime.h>
#include "tefem_peer.cuh"

// Peer-memory reductions (tefem_peer.cuh): available once tefem_comm_shm_alloc / _open have run on every rank.
bool tefem_comm_has_peer(const tefem_comm* c);
int64_t tefem_comm_peer_cap(const tefem_comm* c);
tefem::PeerCtx tefem_comm_peer_ctx(tefem_comm* c, int channel);      // next epoch of that channel

// Peer-mapped vector slots (tefem_comm_vec_alloc / _open): slot k of rank q
bool tefem_comm_has_vectors(const tefem_comm* c);
int tefem_comm_n_slots(const tefem_comm* c);
int64_t tefem_comm_slot_doubles(const tefem_comm* c);
double* tefem_comm_slot(const tefem_comm* c, int rank, int slot);
int tefem_comm_rank(const tefem_comm* c);
// device flag a timed-out peer wait raises (read it with the solver's own asynchronous copies instead of the blocking
// tefem_comm_check), and the error a raised flag maps to
const int* tefem_comm_err_ptr(const tefem_comm* c);
int tefem_comm_report_timeout();
int tefem_comm_world(const tefem_comm* c);

// Packed point-to-point exchange with the neighbouring row blocks: for neighbour k, the doubles
// send[4*send_ptr[k] .. 4*send_ptr[k+1]) go to rank nbr[k] and recv[4*recv_ptr[k] .. 4*recv_ptr[k+1])
// arrive from it, all inside one ncclGroupStart/End.
int tefem_comm_neighbor_exchange(tefem_comm* c, int n_nbr, const int32_t* nbr, const int32_t* send_ptr,
                                 const double* send, const int32_t* recv_ptr, double* recv, cudaStream_t st);
