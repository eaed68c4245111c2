// Coset-sharded commitment trees and owner-based fetches (multi-GPU).
//
// A committed matrix has its level-0 items (row hashes for the trace tree, first-level nodes for the constraint tree) indexed
// i = (k * G + g) * blk + j : rank g owns, for every k < n, the aligned block of blk = 2^log_blk consecutive items.  Each
// rank builds the n complete subtrees over its blocks, the n subtree roots per rank are all-gathered, interleaved into the
// level with n*G nodes, and the upper tree is finished redundantly on every rank (so all ranks hold the same root).
// With G == 1 the same code runs without communication.
#pragma once
#include <array>
#include "common.cuh"

namespace dg {

typedef std::array<uint8_t, 32> Digest;

struct ShardLocation { int owner; bool upper; uint64_t index; };   // upper: index into the replicated upper heap, else local heap / item index

struct ShardGeom {
    uint64_t n;          // number of blocks per rank
    int log_blk;         // items per block
    int log_g;           // log2(world)
    uint64_t items() const { return (n << log_blk) << log_g; }
    // level-0 item i -> owner and index into the owner's local item array ([k][j])
    ShardLocation item(uint64_t i) const;
    // internal node with global heap index h (1 <= h < items())
    ShardLocation node(uint64_t h) const;
};

// one 16-byte unit (or 32-byte item = two consecutive units) to fetch: `index` counts items of `bytes` bytes from `base` on the
// rank `owner` (owner < 0: replicated, every rank reads its own copy)
struct FetchRef { const void *base; uint64_t index; int owner; };

// All device->host fetches of the openings stage in ONE pass: every rank registers the same requests in the same order, one kernel
// gathers the units this rank owns (zeros elsewhere), one all-gather + one copy bring every rank's units to every host, and the host
// picks each unit from its owner.  Replaces ~40 blocking index-upload / gather / download round trips per proof (r01: 1.3-1.6 ms).
class FetchBatch {
public:
    explicit FetchBatch(Context &c) : c_(c) {}
    size_t add16(const FetchRef &r) { return push(r.base, r.index, r.owner); }                       // returns the unit offset of the value
    size_t add32(const FetchRef &r) { size_t o = push(r.base, 2 * r.index, r.owner); push(r.base, 2 * r.index + 1, r.owner); return o; }
    void run();
    fe value(size_t off) const { fe v; memcpy(&v, out_.data() + off * 16, 16); return v; }
    Digest digest(size_t off) const { Digest d; memcpy(d.data(), out_.data() + off * 16, 32); return d; }
    size_t units() const { return owner_.size(); }
private:
    size_t push(const void *base, uint64_t unit, int owner) {
        const bool mine = owner < 0 || owner == c_.rank;
        req_.push_back(mine ? (unsigned long long)(uintptr_t)base : 0ULL);
        req_.push_back(mine ? unit : 0ULL);
        owner_.push_back(owner);
        return owner_.size() - 1;
    }
    Context &c_;
    std::vector<unsigned long long> req_;      // (base, unit) pairs
    std::vector<int> owner_;
    std::vector<uint8_t> out_;
};

struct ShardedTree {
    ShardGeom geom;
    const void *items_local = nullptr;   // n * blk digests, [k][j]
    DevBuf local_nodes;                  // heap over the local items (valid for levels with >= n nodes)
    DevBuf upper;                        // replicated heap: 2 * n * G digests, level with n*G nodes at [nG, 2nG)
    const void *upper_p = nullptr;       // the upper heap (one rank: the local heap itself)
    Digest root;

    // fetch_root = false leaves the root on the device (upper_p + 32): no host synchronisation
    void build(Context &c, const void *items_local_dev, uint64_t n, int log_blk, bool fetch_root = true);
    const void *root_dev() const { return (const uint8_t *)upper_p + 32; }
    // collective fetches (every rank passes the same lists); results in request order
    std::vector<Digest> fetch_nodes(Context &c, const std::vector<uint64_t> &heap_indices) const;
    std::vector<Digest> fetch_items(Context &c, const std::vector<uint64_t> &item_indices) const;
    // the same locations as references for a FetchBatch
    FetchRef item_ref(uint64_t item_index) const { ShardLocation l = geom.item(item_index); return FetchRef{items_local, l.index, l.owner}; }
    FetchRef node_ref(uint64_t heap_index) const {
        ShardLocation l = geom.node(heap_index);
        if (l.upper) return FetchRef{upper_p, l.index, -1};
        return FetchRef{local_nodes.p, l.index, l.owner};
    }
};

// owner-based exchange: every rank has filled `local` (count items of item_bytes) with the entries it owns; returns, for each
// request q, the entry produced by owners[q]
std::vector<uint8_t> exchange_owned(Context &c, const void *d_local, size_t count, size_t item_bytes, const std::vector<int> &owners);

void merkle_build_partial(Context &c, const void *leaves, void *nodes, unsigned long long L, unsigned long long stop);
void merkle_build(Context &c, const void *leaves, void *nodes, unsigned long long L);
void merkle_finish(Context &c, void *nodes, unsigned long long m);
void interleave_roots(Context &c, const void *gathered, void *upper, unsigned long long n, int log_g);
void transpose_cosets(Context &c, const fe *in, fe *out, int log_n, int log_c, int batch);   // [batch][2^log_c][n] -> [batch][n][2^log_c]
void constraint_items_local(Context &c, const fe *evals_local, int log_n, int log_nc, void *items);   // [k][c4_local] digests

}  // namespace dg
