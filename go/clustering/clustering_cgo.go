// Package clustering -- cgo shim over libimageclust_b200.so.
//
// Drop-in replacement for internal/clustering/clustering.go of monahand1023/imageclust:
// the two exported signatures the application uses are kept verbatim
// (PerformClusteringWithConstraints, clustering.go:198; CalculateOptimalClusters,
// clustering.go:168), everything else happens behind the C ABI declared in
// include/imageclust_b200.h.
//
// NOT COMPILED IN THIS REPOSITORY'S BUILD ENVIRONMENT (no Go toolchain in the image);
// the identical C entry points are exercised through the Python ctypes binding
// (imageclust_b200/_lib.py, tests/test_gpu_parity.py).  The application already
// builds with CGO_ENABLED=1 (reference Dockerfile:56, internal/gocv/flags.go:3-5).
package clustering

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../imageclust_b200 -limageclust_b200 -lstdc++ -lm -ldl -lpthread -lrt
#include <stdlib.h>
#include "imageclust_b200.h"
*/
import "C"

import (
	"fmt"
	"log"
	"sync"
	"unsafe"
)

var (
	ctxOnce sync.Once
	ctx     *C.ic_ctx
	ctxErr  error
	ctxMu   sync.Mutex // one clustering at a time per context (the reference function is re-entrant)
)

func engine() (*C.ic_ctx, error) {
	ctxOnce.Do(func() {
		if rc := C.ic_create(&ctx, 0); rc != C.IC_OK {
			ctxErr = fmt.Errorf("ic_create failed (%d): a B200 is required, there is no CPU fallback", int(rc))
		}
	})
	return ctx, ctxErr
}

// CalculateOptimalClusters keeps the reference signature (clustering.go:168).
func CalculateOptimalClusters(totalItems, minSize, maxSize int) (int, error) {
	var out C.int64_t
	switch rc := C.ic_optimal_clusters(C.int64_t(totalItems), C.int64_t(minSize), C.int64_t(maxSize), &out); rc {
	case C.IC_OK:
		return int(out), nil
	case C.IC_ERR_TOO_FEW:
		return 0, fmt.Errorf("total items (%d) less than minimum cluster size (%d)", totalItems, minSize)
	case C.IC_ERR_UNSAT:
		return 0, fmt.Errorf("cannot satisfy cluster size constraints with total items (%d), minSize (%d), and maxSize (%d)", totalItems, minSize, maxSize)
	default:
		return 0, fmt.Errorf("invalid cluster size constraints: total items (%d), minSize (%d), maxSize (%d)", totalItems, minSize, maxSize)
	}
}

// PerformClusteringWithConstraints keeps the reference signature (clustering.go:198).
func PerformClusteringWithConstraints(embeddings [][]float32, productReferenceIDs []string, minSize, maxSize int) (map[int][]string, bool) {
	n := len(embeddings)
	log.Printf("Total items for clustering: %d", n)
	nClusters, err := CalculateOptimalClusters(n, minSize, maxSize)
	if err != nil {
		log.Printf("Clustering constraint error: %v", err)
		return nil, false
	}
	log.Printf("Optimal number of clusters calculated: %d", nClusters)
	if len(productReferenceIDs) < n { // the reference would panic at clustering.go:276
		log.Printf("productReferenceIDs shorter than embeddings (%d < %d)", len(productReferenceIDs), n)
		return nil, false
	}
	d := len(embeddings[0])
	for _, row := range embeddings {
		if len(row) != d { // the reference would panic at clustering.go:149-151
			log.Printf("embeddings have different lengths")
			return nil, false
		}
	}
	c, err := engine()
	if err != nil {
		log.Printf("%v", err)
		return nil, false
	}

	// Go pointers never cross the boundary: flatten into the pinned staging buffer,
	// which is also the source of the single H2D copy.
	count := n * d
	if count == 0 {
		count = 1
	}
	stage := C.ic_pinned_alloc(C.size_t(count * 4))
	if stage == nil {
		log.Printf("ic_pinned_alloc failed")
		return nil, false
	}
	defer C.ic_pinned_free(stage)
	flat := unsafe.Slice((*float32)(stage), count)
	for i, row := range embeddings {
		copy(flat[i*d:(i+1)*d], row)
	}

	offsets := make([]C.int32_t, n+1)
	members := make([]C.int32_t, n+1)
	var k C.int32_t
	var st C.ic_stats

	ctxMu.Lock()
	rc := C.ic_cluster_with_constraints(c, (*C.float)(stage), C.int64_t(n), C.int64_t(d), C.int64_t(d),
		C.int64_t(minSize), C.int64_t(maxSize), &offsets[0], &members[0], &k, &st)
	var msg string
	if rc != C.IC_OK {
		msg = C.GoString(C.ic_last_error(c))
	}
	ctxMu.Unlock()
	if rc != C.IC_OK {
		log.Printf("clustering failed (%d): %s", int(rc), msg)
		return nil, false
	}

	// what workflow.go:89-97 cannot see in the map (SURVEY 8f-2): kept for Report()
	rep := RunReport{Items: n, Clusters: int(k), Merges: int(st.n_merges), NearTies: int(st.n_near_ties),
		NearTieTol: float64(st.near_tie_tol), Exhausted: st.exhausted != 0, ReferenceArithmetic: st.exact != 0,
		PairsReevaluated: int64(st.n_exact), FilterViolations: int(st.n_filter_viol), OrderViolations: int(st.n_order_viol),
		Restarts: int(st.n_restarts)}
	seen := make([]bool, n)
	for i := 0; i < int(offsets[int(k)]); i++ {
		seen[int(members[i])] = true
	}
	for i := 0; i < n; i++ {
		if !seen[i] {
			rep.DroppedItems = append(rep.DroppedItems, productReferenceIDs[i])
		}
	}
	reportMu.Lock()
	lastReport = rep
	reportMu.Unlock()

	clusterMap := make(map[int][]string, int(k))
	for id := 0; id < int(k); id++ {
		lo, hi := int(offsets[id]), int(offsets[id+1])
		refs := make([]string, hi-lo)
		for i := lo; i < hi; i++ {
			refs[i-lo] = productReferenceIDs[int(members[i])]
		}
		clusterMap[id] = refs
	}
	if st.n_near_ties > 0 {
		log.Printf("%d of %d merges had a runner-up within %g relative (near-ties)", int(st.n_near_ties), int(st.n_merges), float32(st.near_tie_tol))
	}
	// what workflow.go:89-97 cannot see in the map: items whose cluster stayed below minSize are absent (clustering.go:268-271)
	if kept := int(offsets[int(k)]); kept < n {
		log.Printf("%d of %d items are in no cluster (their clusters stayed below minSize %d)", n-kept, n, minSize)
	}
	log.Printf("Clustering successful. Formed %d valid clusters.", len(clusterMap))
	return clusterMap, true
}

// RunReport is what PerformClusteringWithConstraints' (map, bool) cannot carry (SURVEY 8f-2, workflow.go:89-97):
// the items that are in no cluster because theirs stayed below minSize (clustering.go:268-271), the near-tie count, and
// the checks behind "same merge sequence as the reference's arithmetic" (both violation counts must be 0).
type RunReport struct {
	Items, Clusters, Merges int
	DroppedItems            []string
	NearTies                int
	NearTieTol              float64
	Exhausted               bool // the loop ended with no admissible pair (clustering.go:222-225)
	ReferenceArithmetic     bool // pairs that could decide a merge were evaluated as WardDistance of two fp32 centroids
	PairsReevaluated        int64
	FilterViolations        int
	OrderViolations         int
	Restarts                int
}

var (
	reportMu   sync.Mutex
	lastReport RunReport
)

// Report returns the report of the LAST successful PerformClusteringWithConstraints call of this process.
func Report() RunReport {
	reportMu.Lock()
	defer reportMu.Unlock()
	return lastReport
}

// Linkage returns the merge trace of the LAST clustering as a dendrogram in the layout of
// scipy.cluster.hierarchy.linkage: rows {keyLo, keyHi, sqrt(2*dist), size}; keys are item indices or n + t for the
// cluster made by merge t.  The reference has no persistent form of its result (workflow.go:99 keeps the map only);
// a caller that stores the linkage can re-cut it at other minSize / maxSize without clustering again.
func Linkage(n int) ([][4]float64, bool) {
	c, err := engine()
	if err != nil || n <= 0 {
		return nil, false
	}
	z := make([]C.double, 4*n)
	var rows C.int64_t
	ctxMu.Lock()
	rc := C.ic_get_linkage(c, &z[0], C.int64_t(n), &rows)
	ctxMu.Unlock()
	if rc != C.IC_OK {
		return nil, false
	}
	out := make([][4]float64, int(rows))
	for t := range out {
		for j := 0; j < 4; j++ {
			out[t][j] = float64(z[4*t+j])
		}
	}
	return out, true
}
