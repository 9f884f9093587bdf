// Host-side structure pipeline (ordering -> etree -> postorder -> column counts ->
// supernodes -> symbolic factorisation -> matrix input).
//
// In the reference all of this is Julia host code that STAYS on the host
// (north_star); there is no Julia in this image, so the host mirror of the
// reference interface (sparspak.jl_b200/*.py) calls these C++ restatements to
// produce exactly the flat 1-based Int64 arrays the numeric C-ABI consumes.
// Every array crossing this API is 1-based int64, laid out as in the reference
// structs (_SparseBase, SpkSparseBase.jl:99-125), so golden vectors from the
// reference's tests compare directly.
//
// Behavioural sources (reference file:line) are cited per function.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>
#include <limits>

typedef int64_t I;

#define SPKH_API extern "C" __attribute__((visibility("default")))

// ---------------------------------------------------------------------------
// Multiple minimum degree ordering (SpkMMD.jl:78-633).  Quotient-graph MMD with
// multiple elimination (delta = 0, SpkMMD.jl:79), external degree and
// indistinguishable-node merging.  State lives in one struct; arrays are used
// 1-based (index 0 unused) so node ids match the reference's.
// ---------------------------------------------------------------------------
namespace {

struct Mmd {
    I n;
    const I* xadj;           // 1-based, length n+1
    std::vector<I> adjncy;   // private copy, 1-based
    std::vector<I> deghead;  // degree 0..n-1  -> index deg (0-based degree)
    std::vector<I> degnext, degprev, supersize, elimnext, marker, mergeparent, needsupdate;
    I* invp;                 // 1-based view
    I maxint;

    void eliminate(I mdnode, I tag);
    void update(I elimhead, I delta, I& mindeg, I& tag);
    void place(I deg_in, I& mindeg, I enode);
};

// SpkMMD.jl:239-353
void Mmd::eliminate(I mdnode, I tag) {
    marker[mdnode] = tag;
    I elmnt = 0;
    I rloc = xadj[mdnode], rlmt = xadj[mdnode + 1] - 1;
    for (I i = xadj[mdnode]; i <= xadj[mdnode + 1] - 1; ++i) {
        I nb = adjncy[i];
        if (nb == 0) break;
        if (marker[nb] < tag) {
            marker[nb] = tag;
            if (invp[nb] == 0) { adjncy[rloc++] = nb; }
            else { elimnext[nb] = elmnt; elmnt = nb; }
        }
    }
    // absorb the reach sets of adjacent elements
    while (elmnt > 0) {
        adjncy[rlmt] = -elmnt;
        I j = xadj[elmnt], jstop = xadj[elmnt + 1];
        I node = adjncy[j];
        while (node != 0) {
            if (node < 0) { j = xadj[-node]; jstop = xadj[-node + 1]; }
            else {
                if (marker[node] < tag && degnext[node] >= 0) {
                    marker[node] = tag;
                    while (rloc >= rlmt) {           // borrow storage of eliminated nodes
                        I link = -adjncy[rlmt];
                        rloc = xadj[link]; rlmt = xadj[link + 1] - 1;
                    }
                    adjncy[rloc++] = node;
                }
                ++j;
            }
            if (j >= jstop) break;
            node = adjncy[j];
        }
        elmnt = elimnext[elmnt];
    }
    if (rloc <= rlmt) adjncy[rloc] = 0;
    // visit every node of the reach set
    I i = xadj[mdnode], istop = xadj[mdnode + 1];
    I rnode = adjncy[i];
    while (rnode != 0) {
        if (rnode < 0) { i = xadj[-rnode]; istop = xadj[-rnode + 1]; }
        else {
            I pv = degprev[rnode];
            if (pv != 0) {                           // unlink from the degree lists
                I nx = degnext[rnode];
                if (nx > 0) degprev[nx] = pv;
                if (pv > 0) degnext[pv] = nx; else deghead[-pv] = nx;
            }
            I xq = xadj[rnode];
            for (I j = xadj[rnode]; j <= xadj[rnode + 1] - 1; ++j) {
                I nb = adjncy[j];
                if (nb == 0) break;
                if (marker[nb] < tag) adjncy[xq++] = nb;
            }
            I nq = xq - xadj[rnode];
            if (nq <= 0) {                           // indistinguishable from mdnode: merge
                supersize[mdnode] += supersize[rnode];
                supersize[rnode] = 0; mergeparent[rnode] = mdnode;
                marker[rnode] = maxint;
            } else {
                needsupdate[rnode] = nq + 1;
                adjncy[xq++] = mdnode;
                if (xq < xadj[rnode + 1]) adjncy[xq] = 0;
            }
            degprev[rnode] = 0; ++i;
        }
        if (i >= istop) break;
        rnode = adjncy[i];
    }
}

// SpkMMD.jl:562-569
void Mmd::place(I deg, I& mindeg, I enode) {
    deg -= supersize[enode];
    I first = deghead[deg];
    deghead[deg] = enode; degnext[enode] = first;
    degprev[enode] = -deg; needsupdate[enode] = 0;
    if (first > 0) degprev[first] = enode;
    if (deg < mindeg) mindeg = deg;
}

// SpkMMD.jl:384-559
void Mmd::update(I elimhead, I delta, I& mindeg, I& tag) {
    I mindeglimit = mindeg + delta;
    I elimnode = elimhead;
    while (elimnode > 0) {
        I mtag = tag + mindeglimit;
        if (mtag >= maxint) {
            tag = 1; mtag = tag + mindeglimit;
            for (I v = 1; v <= n; ++v) if (marker[v] < maxint) marker[v] = 0;
        }
        I q2head = 0, qxhead = 0, elimsize = 0;
        I i = xadj[elimnode], istop = xadj[elimnode + 1];
        I enode = adjncy[i];
        while (enode != 0) {
            if (enode < 0) { i = xadj[-enode]; istop = xadj[-enode + 1]; }
            else {
                if (supersize[enode] != 0) {
                    elimsize += supersize[enode];
                    marker[enode] = mtag;
                    if (needsupdate[enode] > 0) {
                        if (needsupdate[enode] != 2) { elimnext[enode] = qxhead; qxhead = enode; }
                        else { elimnext[enode] = q2head; q2head = enode; }
                    }
                }
                ++i;
            }
            if (i >= istop) break;
            enode = adjncy[i];
        }
        // nodes adjacent to exactly two elements
        enode = q2head;
        while (enode > 0) {
            if (needsupdate[enode] > 0) {
                ++tag; I deg = elimsize;
                I istart = xadj[enode];
                I nb = adjncy[istart];
                if (nb == elimnode) nb = adjncy[istart + 1];
                if (invp[nb] == 0) deg += supersize[nb];
                else {
                    I ii = xadj[nb], iistop = xadj[nb + 1];
                    I node = adjncy[ii];
                    while (node != 0) {
                        if (node < 0) { ii = xadj[-node]; iistop = xadj[-node + 1]; }
                        else {
                            if (node != enode && supersize[node] != 0) {
                                if (marker[node] < tag) { marker[node] = tag; deg += supersize[node]; }
                                else if (needsupdate[node] > 0) {
                                    if (needsupdate[node] == 2) {
                                        supersize[enode] += supersize[node];
                                        supersize[node] = 0; marker[node] = maxint;
                                        mergeparent[node] = enode;
                                    }
                                    needsupdate[node] = 0; degprev[node] = 0;
                                }
                            }
                            ++ii;
                        }
                        if (ii >= iistop) break;
                        node = adjncy[ii];
                    }
                }
                place(deg, mindeg, enode);
            }
            enode = elimnext[enode];
        }
        // nodes adjacent to more than two elements
        enode = qxhead;
        while (enode > 0) {
            if (needsupdate[enode] > 0) {
                ++tag; I deg = elimsize;
                for (I ii = xadj[enode]; ii <= xadj[enode + 1] - 1; ++ii) {
                    I nb = adjncy[ii];
                    if (nb == 0) break;
                    if (marker[nb] < tag) {
                        marker[nb] = tag;
                        if (invp[nb] == 0) deg += supersize[nb];
                        else {
                            I j = xadj[nb], jstop = xadj[nb + 1];
                            I node = adjncy[j];
                            while (node != 0) {
                                if (node < 0) { j = xadj[-node]; jstop = xadj[-node + 1]; }
                                else {
                                    if (marker[node] < tag) { marker[node] = tag; deg += supersize[node]; }
                                    ++j;
                                }
                                if (j >= jstop) break;
                                node = adjncy[j];
                            }
                        }
                    }
                }
                place(deg, mindeg, enode);
            }
            enode = elimnext[enode];
        }
        tag = mtag; elimnode = elimnext[elimnode];
    }
}

} // namespace

// mmd!(g, order)  — SpkMMD.jl:38-42, 78-209, 588-633.
// xadj[n+1], adj[xadj[n]-1] 1-based values in 0-based C arrays; perm/invp length n, 1-based values.
SPKH_API int spkh_mmd(I n, const I* xadj0, const I* adj0, I* perm0, I* invp0) {
    if (n <= 0) return 0;
    Mmd m;
    m.n = n; m.maxint = std::numeric_limits<I>::max();
    std::vector<I> xadj(n + 2);
    for (I i = 1; i <= n + 1; ++i) xadj[i] = xadj0[i - 1];
    m.xadj = xadj.data();
    I ne = xadj[n + 1] - 1;
    m.adjncy.assign(ne + 2, 0);
    for (I k = 1; k <= ne; ++k) m.adjncy[k] = adj0[k - 1];
    m.deghead.assign(n + 1, 0);
    m.degnext.assign(n + 1, 0); m.degprev.assign(n + 1, 0);
    m.supersize.assign(n + 1, 1); m.elimnext.assign(n + 1, 0);
    m.marker.assign(n + 1, 0); m.mergeparent.assign(n + 1, 0); m.needsupdate.assign(n + 1, 0);
    std::vector<I> invp(n + 1, 0);
    m.invp = invp.data();
    const I delta = 0;

    for (I node = 1; node <= n; ++node) {
        I ndeg = xadj[node + 1] - xadj[node];
        I f = m.deghead[ndeg];
        m.deghead[ndeg] = node; m.degnext[node] = f;
        if (f > 0) m.degprev[f] = node;
        m.degprev[node] = -ndeg;
    }
    I num = 1;
    for (I md = m.deghead[0]; md > 0; md = m.degnext[md]) {
        m.marker[md] = m.maxint; invp[md] = num++;
    }
    m.deghead[0] = 0;
    I tag = 1, mindeg = 1;
    bool done = false;
    while (num <= n && !done) {
        while (m.deghead[mindeg] <= 0) ++mindeg;
        I mindeglimit = mindeg + delta;
        if (delta < 0) mindeglimit = mindeg;
        I elimhead = 0;
        for (;;) {
            I md = m.deghead[mindeg];
            bool pass = false;
            while (md <= 0) {
                ++mindeg;
                if (mindeg > mindeglimit) { pass = true; break; }
                md = m.deghead[mindeg];
            }
            if (pass) break;
            I nx = m.degnext[md];
            m.deghead[mindeg] = nx;
            if (nx > 0) m.degprev[nx] = -mindeg;
            invp[md] = num;
            if (num + m.supersize[md] > n) { done = true; break; }
            ++tag;
            if (tag >= m.maxint) {
                tag = 1;
                for (I v = 1; v <= n; ++v) if (m.marker[v] < m.maxint) m.marker[v] = 0;
            }
            m.eliminate(md, tag);
            num += m.supersize[md];
            m.elimnext[md] = elimhead; elimhead = md;
        }
        if (done || num > n) break;
        m.update(elimhead, delta, mindeg, tag);
    }
    // final numbering through the merge forest (SpkMMD.jl:588-633)
    std::vector<I> lastnum(n + 1, 0);
    for (I v = 1; v <= n; ++v) if (m.mergeparent[v] == 0) lastnum[v] = invp[v];
    for (I node = 1; node <= n; ++node) {
        I parent = m.mergeparent[node];
        if (parent > 0) {
            I root = 0;
            while (parent > 0) { root = parent; parent = m.mergeparent[parent]; }
            I k = lastnum[root] + 1;
            invp[node] = k; lastnum[root] = k;
            I v = node;
            while (v != root) { I p = m.mergeparent[v]; m.mergeparent[v] = root; v = p; }
        }
    }
    for (I v = 1; v <= n; ++v) { invp0[v - 1] = invp[v]; perm0[invp[v] - 1] = v; }
    return 0;
}

// ---------------------------------------------------------------------------
// Geometric nested dissection for an nx*ny*nz grid with `dof` unknowns per
// node.  NOT in the reference: supplied through the
// reference's ordering-callback seam findorder!(s, orderfunction)
// (SpkSparseSolver.jl:103-111).  Node (x,y,z) has scalar index
// ((z*ny + y)*nx + x)*dof + d + 1 (x fastest).  Recursion: cut the longest
// axis in the middle plane, number both halves first, the separator last;
// boxes with <= leaf nodes are numbered lexicographically.
// ---------------------------------------------------------------------------
namespace {
struct NdCtx { I nx, ny, nz, dof, leaf; I* perm; I next; };

inline void nd_node(NdCtx& c, I x, I y, I z) {
    for (I d = 0; d < c.dof; ++d) c.perm[c.next++] = ((z * c.ny + y) * c.nx + x) * c.dof + d + 1;
}
void nd_emit(NdCtx& c, I x0, I x1, I y0, I y1, I z0, I z1) {
    for (I z = z0; z < z1; ++z) for (I y = y0; y < y1; ++y) for (I x = x0; x < x1; ++x) nd_node(c, x, y, z);
}
// Separator plane: nodes that touch a face of the box lying on an ANCESTOR separator (any box face
// that is not on the domain boundary) are numbered last.  Each of them brings one new row into the
// structure when it is eliminated, so numbering them last keeps the rest of the plane one
// fundamental supernode instead of breaking it every few columns.
void nd_emit_sep(NdCtx& c, I x0, I x1, I y0, I y1, I z0, I z1, I bx0, I bx1, I by0, I by1, I bz0, I bz1) {
    auto touches = [&](I x, I y, I z) {
        return (x == bx0 && bx0 > 0) || (x == bx1 - 1 && bx1 < c.nx) || (y == by0 && by0 > 0) ||
               (y == by1 - 1 && by1 < c.ny) || (z == bz0 && bz0 > 0) || (z == bz1 - 1 && bz1 < c.nz);
    };
    for (int pass = 0; pass < 2; ++pass)
        for (I z = z0; z < z1; ++z) for (I y = y0; y < y1; ++y) for (I x = x0; x < x1; ++x)
            if ((int)touches(x, y, z) == pass) nd_node(c, x, y, z);
}
void nd_rec(NdCtx& c, I x0, I x1, I y0, I y1, I z0, I z1) {
    I lx = x1 - x0, ly = y1 - y0, lz = z1 - z0;
    if (lx <= 0 || ly <= 0 || lz <= 0) return;
    if (lx * ly * lz <= c.leaf || (lx <= 2 && ly <= 2 && lz <= 2)) { nd_emit(c, x0, x1, y0, y1, z0, z1); return; }
    if (lx >= ly && lx >= lz) {
        I m = x0 + lx / 2;
        nd_rec(c, x0, m, y0, y1, z0, z1); nd_rec(c, m + 1, x1, y0, y1, z0, z1);
        nd_emit_sep(c, m, m + 1, y0, y1, z0, z1, x0, x1, y0, y1, z0, z1);
    } else if (ly >= lz) {
        I m = y0 + ly / 2;
        nd_rec(c, x0, x1, y0, m, z0, z1); nd_rec(c, x0, x1, m + 1, y1, z0, z1);
        nd_emit_sep(c, x0, x1, m, m + 1, z0, z1, x0, x1, y0, y1, z0, z1);
    } else {
        I m = z0 + lz / 2;
        nd_rec(c, x0, x1, y0, y1, z0, m); nd_rec(c, x0, x1, y0, y1, m + 1, z1);
        nd_emit_sep(c, x0, x1, y0, y1, m, m + 1, x0, x1, y0, y1, z0, z1);
    }
}
} // namespace

SPKH_API int spkh_nd_grid(I nx, I ny, I nz, I dof, I leaf, I* perm, I* invp) {
    NdCtx c{nx, ny, nz, dof, leaf, perm, 0};
    nd_rec(c, 0, nx, 0, ny, 0, nz);
    I n = nx * ny * nz * dof;
    if (c.next != n) return -1;
    for (I k = 0; k < n; ++k) invp[perm[k] - 1] = k + 1;
    return 0;
}

// ---------------------------------------------------------------------------
// Elimination tree with path compression (SpkETree.jl:62-88).
// ---------------------------------------------------------------------------
SPKH_API int spkh_etree(I n, const I* xadj, const I* adj, const I* rperm, const I* rinvp, I* parent) {
    std::vector<I> anc(n + 1, 0);
    for (I i = 1; i <= n; ++i) {
        parent[i - 1] = 0; anc[i] = 0;
        I v = rperm[i - 1];
        for (I j = xadj[v - 1]; j <= xadj[v] - 1; ++j) {
            I nbr = rinvp[adj[j - 1] - 1];
            if (nbr < i) {
                while (anc[nbr] != 0 && anc[nbr] != i) { I nx = anc[nbr]; anc[nbr] = i; nbr = nx; }
                if (anc[nbr] == 0) { parent[nbr - 1] = i; anc[nbr] = i; }
            }
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------
// Postordering of the etree, optionally weighted so the heaviest child comes
// last (SpkETree.jl:106-139 drivers; :153-174 first-son/brother form; :239-274
// weighted form; :190-220 the stack traversal + parent relabel).
// parent, rperm, rinvp (and weight if given) are permuted in place.
// ---------------------------------------------------------------------------
SPKH_API int spkh_postorder(I n, I* parent0, I* rperm0, I* rinvp0, I* weight0) {
    if (n <= 0) return 0;
    std::vector<I> fson(n + 1, 0), bro(n + 1, 0), lson(n + 1, 0), stack(n + 1, 0), invpos(n + 1, 0);
    I* parent = parent0 - 1; I* w = weight0 ? weight0 - 1 : nullptr;
    I lroot = n;
    for (I v = n - 1; v >= 1; --v) {
        I p = parent[v];
        if (p <= 0 || p == v) { bro[lroot] = v; lroot = v; continue; }
        if (!w) { bro[v] = fson[p]; fson[p] = v; continue; }
        I ls = lson[p];
        if (ls != 0) {
            if (w[v] >= w[ls]) { bro[v] = fson[p]; fson[p] = v; }
            else { bro[ls] = v; lson[p] = v; }
        } else { fson[p] = v; lson[p] = v; }
    }
    bro[lroot] = 0;
    // depth-first traversal: number a vertex when popped
    I num = 0, top = 0, v = n;
    while (v > 0) {
        while (v > 0) { stack[++top] = v; v = fson[v]; }
        while (v <= 0 && top > 0) {
            v = stack[top--];
            invpos[v] = ++num;
            v = bro[v];
        }
    }
    for (I u = 1; u <= num; ++u) {
        I nu = invpos[u], p = parent[u];
        if (p > 0) p = invpos[p];
        bro[nu] = p;
    }
    for (I u = 1; u <= n; ++u) parent[u] = bro[u];
    if (w) {
        for (I u = 1; u <= n; ++u) stack[invpos[u]] = w[u];
        for (I u = 1; u <= n; ++u) w[u] = stack[u];
    }
    for (I u = 1; u <= n; ++u) stack[u] = invpos[rinvp0[u - 1]];
    for (I u = 1; u <= n; ++u) rinvp0[u - 1] = stack[u];
    for (I u = 1; u <= n; ++u) rperm0[rinvp0[u - 1] - 1] = u;
    return 0;
}

// ---------------------------------------------------------------------------
// Column counts by row-subtree traversal (SpkSymFct.jl:37-59).  Returns nnz(L).
// ---------------------------------------------------------------------------
SPKH_API I spkh_colcounts(I n, const I* xadj, const I* adj, const I* perm, const I* invp,
                          const I* parent, I* colcnt) {
    std::vector<I> marker(n + 1, 0);
    I nlnz = n;
    for (I i = 1; i <= n; ++i) {
        marker[i] = i; colcnt[i - 1] = 1;
        I v = perm[i - 1];
        for (I k = xadj[v - 1]; k <= xadj[v] - 1; ++k) {
            I j = invp[adj[k - 1] - 1];
            if (j < i) {
                while (marker[j] != i) {
                    ++colcnt[j - 1]; ++nlnz;
                    marker[j] = i; j = parent[j - 1];
                }
            }
        }
    }
    return nlnz;
}

// ---------------------------------------------------------------------------
// Fundamental supernodes, split near maxsize with the reference's FLOATING
// POINT rule (SpkSymFct.jl:415-464): widths may exceed
// maxsize.  xsuper must have room for n+1 entries.  out[0]=nsuper, out[1]=nsub.
// ---------------------------------------------------------------------------
SPKH_API int spkh_findsupernodes(I n, const I* parent, const I* colcnt, I maxsize,
                                 I* xsuper, I* snode, I* out) {
    std::vector<I> marker(n + 2, 0);
    I nsuper = 1; xsuper[0] = 1;
    for (I k = 2; k <= n; ++k) {
        if (parent[k - 2] == k && colcnt[k - 2] == colcnt[k - 1] + 1) continue;
        xsuper[nsuper++] = k;
    }
    xsuper[nsuper] = n + 1;
    for (I js = 1; js <= nsuper; ++js) {
        I first = xsuper[js - 1], nextfirst = xsuper[js];
        I sz = nextfirst - first;
        if (sz > maxsize) {
            double delta = (double)sz / (1.0 + (double)sz / (double)maxsize);
            I limit = (I)std::floor((double)nextfirst - delta / 2.0);
            I k = 1;
            I idx = (I)std::floor((double)first + delta);
            while (idx < limit) {
                marker[idx] = 1;
                ++k; idx = (I)std::floor((double)first + (double)k * delta);
            }
        }
    }
    nsuper = 1; snode[0] = 1; I nofsub = colcnt[0];
    for (I k = 2; k <= n; ++k) {
        if (marker[k] != 1 && parent[k - 2] == k && colcnt[k - 2] == colcnt[k - 1] + 1) {
            snode[k - 1] = nsuper; continue;
        }
        ++nsuper; snode[k - 1] = nsuper;
        xsuper[nsuper - 1] = k; nofsub += colcnt[k - 1];
    }
    xsuper[nsuper] = n + 1;
    out[0] = nsuper; out[1] = nofsub;
    return 0;
}

// ---------------------------------------------------------------------------
// Column pointers of the rectangular supernode storage
// (LU: SpkSparseBase.jl:253-284; SPD: SpkSparseSpdBase.jl:359-375).
// xunz may be NULL (SPD).
// ---------------------------------------------------------------------------
SPKH_API int spkh_nonzeroindexs(I n, const I* colcnt, I nsuper, const I* xsuper, I* xlnz, I* xunz) {
    I point = 1, upoint = 1;
    for (I ks = 1; ks <= nsuper; ++ks) {
        I f = xsuper[ks - 1], l = xsuper[ks] - 1, width = l - f + 1;
        for (I j = f; j <= l; ++j) {
            xlnz[j - 1] = point; point += colcnt[f - 1];
            if (xunz) { xunz[j - 1] = upoint; upoint += colcnt[f - 1] - width; }
        }
    }
    xlnz[n] = point;
    if (xunz) xunz[n] = upoint;
    return 0;
}

// ---------------------------------------------------------------------------
// Supernodal symbolic factorisation (SpkSymFct.jl:105-246): the row index list
// of each supernode = sorted merge of its children's lists (minus their own
// columns) and the structure of A(*,first column).  Returns 0, or -1 on the
// reference's "Inconsistency in data structure" condition.
// ---------------------------------------------------------------------------
SPKH_API int spkh_symbolicfact(I n, const I* xadj, const I* adj, const I* perm, const I* invp,
                               const I* colcnt, I nsuper, const I* xsuper, const I* snode,
                               I* xlindx, I* lindx) {
    std::vector<I> marker(n + 1, 0), mrglnk(nsuper + 1, 0), rch(n + 2, 0);
    const I head = 0, tail = n + 1;
    I nzend = 0, point = 1;
    for (I ks = 1; ks <= nsuper; ++ks) { xlindx[ks - 1] = point; point += colcnt[xsuper[ks - 1] - 1]; }
    xlindx[nsuper] = point;
    for (I ks = 1; ks <= nsuper; ++ks) {
        I fst = xsuper[ks - 1], lst = xsuper[ks] - 1;
        I width = lst - fst + 1, len = colcnt[fst - 1];
        I knz = 0; rch[head] = tail;
        I js = mrglnk[ks];
        if (js > 0) {
            // first child: copy its below-block indices (already sorted)
            I jw = xsuper[js] - xsuper[js - 1];
            I b = xlindx[js - 1] + jw, e = xlindx[js] - 1;
            for (I p = e; p >= b; --p) {
                I r = lindx[p - 1]; ++knz;
                marker[r] = ks; rch[r] = rch[head]; rch[head] = r;
            }
            js = mrglnk[js];
            while (js != 0 && knz < len) {           // merge the remaining children
                jw = xsuper[js] - xsuper[js - 1];
                b = xlindx[js - 1] + jw; e = xlindx[js] - 1;
                I nexti = head;
                for (I p = b; p <= e; ++p) {
                    I r = lindx[p - 1];
                    I i = nexti; nexti = rch[i];
                    while (r > nexti) { i = nexti; nexti = rch[i]; }
                    if (r < nexti) { ++knz; rch[i] = r; rch[r] = nexti; marker[r] = ks; nexti = r; }
                }
                js = mrglnk[js];
            }
        }
        if (knz < len) {                             // structure of A(*, fst)
            I node = perm[fst - 1];
            for (I p = xadj[node - 1]; p <= xadj[node] - 1; ++p) {
                I r = invp[adj[p - 1] - 1];
                if (r > fst && marker[r] != ks) {
                    I nexti = head, i = nexti; nexti = rch[i];
                    while (r > nexti) { i = nexti; nexti = rch[i]; }
                    ++knz; rch[i] = r; rch[r] = nexti; marker[r] = ks;
                }
            }
        }
        if (rch[head] != fst) { rch[fst] = rch[head]; rch[head] = fst; ++knz; }
        I nzbeg = nzend + 1; nzend += knz;
        if (nzend + 1 != xlindx[ks]) return -1;
        I i = head;
        for (I p = nzbeg; p <= nzend; ++p) { i = rch[i]; lindx[p - 1] = i; }
        if (len > width) {
            I pcol = lindx[xlindx[ks - 1] + width - 1];
            I ps = snode[pcol - 1];
            mrglnk[ks] = mrglnk[ps]; mrglnk[ps] = ks;
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------
// Matrix input.  The reference scatters A's entries into lnz/unz with a linear
// search per entry (LU: SpkSparseBase.jl:302-372, CSC twin
// SparseCSCInterface.jl:101-169; SPD: SpkSparseSpdBase.jl:234-311).  Here the
// search result (the 1-based destination slot of every stored entry) is
// produced as an index map so it can be reused by every refactorisation and
// handed to the device scatter kernel.
//   dest[k] > 0 : slot in lnz;   dest[k] < 0 : slot -dest[k] in unz;
//   dest[k] == 0: entry ignored (SPD: strict upper triangle of the input).
// colptr/rowval are 1-based CSC.  Returns 0, or k+1 (1-based entry number) of
// the first entry that has "No space for matrix element".
// ---------------------------------------------------------------------------
namespace {
// position of `row` in the sorted index list lindx[b..e] (1-based positions), or -1
inline I find_row(const I* lindx, I b, I e, I row) {
    const I* lo = std::lower_bound(lindx + (b - 1), lindx + e, row);
    if (lo == lindx + e || *lo != row) return -1;
    return (I)(lo - lindx) + 1;
}
}

SPKH_API I spkh_inmatrix_map_lu(I n, const I* colptr, const I* rowval, const I* rinvp, const I* cinvp,
                                const I* snode, const I* xsuper, const I* xlindx, const I* lindx,
                                const I* xlnz, const I* xunz, I* dest) {
    for (I i = 1; i <= n; ++i) {
        for (I p = colptr[i - 1]; p <= colptr[i] - 1; ++p) {
            I inew = rinvp[rowval[p - 1] - 1], jnew = cinvp[i - 1];
            if (inew >= xsuper[snode[jnew - 1] - 1]) {               // lies in L (incl. diagonal block)
                I js = snode[jnew - 1];
                I b = xlindx[js - 1], e = xlindx[js] - 1;
                I pos = find_row(lindx, b, e, inew);
                if (pos < 0) return p;
                dest[p - 1] = xlnz[jnew - 1] + (pos - b);
            } else {                                                  // lies in U (stored by rows)
                I js = snode[inew - 1];
                I width = xsuper[js] - xsuper[js - 1];
                I b = xlindx[js - 1] + width, e = xlindx[js] - 1;
                I pos = find_row(lindx, b, e, jnew);
                if (pos < 0) return p;
                dest[p - 1] = -(xunz[inew - 1] + (pos - b));
            }
        }
    }
    return 0;
}

SPKH_API I spkh_inmatrix_map_spd(I n, const I* colptr, const I* rowval, const I* rinvp, const I* cinvp,
                                 const I* snode, const I* xsuper, const I* xlindx, const I* lindx,
                                 const I* xlnz, I* dest) {
    for (I i = 1; i <= n; ++i) {
        for (I p = colptr[i - 1]; p <= colptr[i] - 1; ++p) {
            I r = rowval[p - 1];
            if (r < i) { dest[p - 1] = 0; continue; }                 // only the lower triangle is used
            I inew = rinvp[r - 1], jnew = cinvp[i - 1];
            if (inew < jnew) std::swap(inew, jnew);
            I js = snode[jnew - 1], fstcol = xsuper[js - 1];
            I b = xlindx[js - 1] + (jnew - fstcol), e = xlindx[js] - 1;
            I pos = find_row(lindx, b, e, inew);
            if (pos < 0) return p;
            dest[p - 1] = xlnz[jnew - 1] + (pos - b) + (jnew - fstcol);
        }
    }
    return 0;
}

// lnz/unz must be zeroed by the caller; values are ADDED (duplicates sum), as in the reference.
SPKH_API int spkh_scatter_values(I nnz, const I* dest, const double* nzval, double* lnz, double* unz) {
    for (I k = 0; k < nnz; ++k) {
        I d = dest[k];
        if (d > 0) lnz[d - 1] += nzval[k];
        else if (d < 0) unz[-d - 1] += nzval[k];
    }
    return 0;
}

// Structure-only work counts: out[0]=sum cc, out[1]=sum cc^2 (as doubles).
SPKH_API int spkh_workcounts(I n, const I* colcnt_per_col, double* out) {
    double s1 = 0, s2 = 0;
    for (I j = 0; j < n; ++j) { double c = (double)colcnt_per_col[j]; s1 += c; s2 += c * c; }
    out[0] = s1; out[1] = s2;
    return 0;
}
