// ccref_harness_test.go — runs the UNMODIFIED reference (sigs.k8s.io/cluster-capacity pkg/framework) on the snapshot files
// this repo's tests use, and dumps what the reference decides, so that the CPU oracles (oracle/ccsim_oracle.c,
// oracle/objref.py) and through them the CUDA path can be pinned against a real run of the reference.
//
// TEST INFRASTRUCTURE. It is not compiled here (the build image has no Go toolchain) and never shipped in the product.
// Recipe (any box with Go >= 1.24 and the reference checkout; no network, the reference vendors its module graph):
//
//	make -C oracle ref REFERENCE=/path/to/cluster-capacity        # see oracle/Makefile
//
// which copies this file into $REFERENCE/pkg/framework/ (a scratch copy when the tree is read-only), runs
//
//	CCREF_CASES=<repo>/tests/golden/ref_cases CCREF_OUT=<repo>/oracle/_ref go test -mod=vendor -run TestCCRefHarness ./pkg/framework/
//
// and leaves one <case>.out.json per case under oracle/_ref/ (git-ignored; it travels to the GPU box with the snapshot).
// tests/test_reference_outputs.py consumes those files when they exist.
//
// Modelled on the reference's own test (pkg/framework/simulator_test.go:200-240): fake clientset loaded with the case's
// Node / Pod objects, utils.BuildKubeSchedulerCompletedConfig, framework.New, SyncWithClient, Run, Report. The scheduler
// configuration is the defaulted v1 configuration with percentageOfNodesToScore: 100 — the canonical mode this repo
// reproduces (every node filtered every cycle, nextStartNodeIndex fixed: schedule_one.go:538-539,697-723). Ties in
// selectHost are random in the reference (schedule_one.go:916), so a consumer compares the order-independent parts:
// instance count, fail type and message, and per-node counts where they are forced (KA5-class cases).
package framework

import (
	"encoding/json"
	"os"
	"path/filepath"
	"sort"
	"strings"
	"testing"
	"time"

	v1 "k8s.io/api/core/v1"
	"k8s.io/apimachinery/pkg/runtime"
	fakeclientset "k8s.io/client-go/kubernetes/fake"
	configv1alpha1 "k8s.io/component-base/config/v1alpha1"
	kubeschedulerconfigv1 "k8s.io/kube-scheduler/config/v1"
	kubeschedulerconfig "k8s.io/kubernetes/pkg/scheduler/apis/config"
	kubeschedulerscheme "k8s.io/kubernetes/pkg/scheduler/apis/config/scheme"

	"sigs.k8s.io/cluster-capacity/pkg/utils"
)

type ccrefCase struct {
	Name     string    `json:"name"`
	Nodes    []v1.Node `json:"nodes"`
	Pods     []v1.Pod  `json:"pods"`
	Template v1.Pod    `json:"template"`
	MaxPods  int       `json:"max_pods"`
	Exclude  []string  `json:"exclude_nodes"`
	// percentageOfNodesToScore for this case: 100 (canonical) unless the case says otherwise (0 = the adaptive default)
	Percentage *int32 `json:"percentageOfNodesToScore"`
}

type ccrefNodeCount struct {
	NodeName string `json:"nodeName"`
	Replicas int    `json:"replicas"`
}

type ccrefOut struct {
	Name         string           `json:"name"`
	Replicas     int32            `json:"replicas"`
	FailType     string           `json:"failType"`
	FailMessage  string           `json:"failMessage"`
	PerNode      []ccrefNodeCount `json:"perNode"`      // sorted by node name
	Sequence     []string         `json:"sequence"`     // node of pod k (order depends on the reference's random tie-breaking)
	RunSeconds   float64          `json:"runSeconds"`   // wall time of Run() (includes the reference's fixed 100 ms start-up sleep)
	SyncSeconds  float64          `json:"syncSeconds"`
	Percentage   int32            `json:"percentageOfNodesToScore"`
	GoMaxProcs   int              `json:"gomaxprocs"`
	NodesInCase  int              `json:"nodes"`
	PodsInCase   int              `json:"pods"`
	ErrorMessage string           `json:"error,omitempty"`
}

func ccrefConfig(pct int32) (*kubeschedulerconfig.KubeSchedulerConfiguration, error) {
	kcfg := &kubeschedulerconfig.KubeSchedulerConfiguration{}
	versioned := kubeschedulerconfigv1.KubeSchedulerConfiguration{}
	versioned.DebuggingConfiguration = *configv1alpha1.NewRecommendedDebuggingConfiguration()
	versioned.PercentageOfNodesToScore = &pct
	kubeschedulerscheme.Scheme.Default(&versioned)
	if err := kubeschedulerscheme.Scheme.Convert(&versioned, kcfg, nil); err != nil {
		return nil, err
	}
	return kcfg, nil
}

func ccrefRun(c *ccrefCase) (out ccrefOut) {
	out.Name = c.Name
	out.NodesInCase, out.PodsInCase = len(c.Nodes), len(c.Pods)
	pct := int32(100)
	if c.Percentage != nil {
		pct = *c.Percentage
	}
	out.Percentage = pct
	kcfg, err := ccrefConfig(pct)
	if err != nil {
		out.ErrorMessage = err.Error()
		return
	}
	completed, err := utils.BuildKubeSchedulerCompletedConfig(kcfg)
	if err != nil {
		out.ErrorMessage = err.Error()
		return
	}
	var objs []runtime.Object
	for i := range c.Nodes {
		objs = append(objs, &c.Nodes[i])
	}
	for i := range c.Pods {
		objs = append(objs, &c.Pods[i])
	}
	client := fakeclientset.NewSimpleClientset(objs...)
	tmpl := c.Template.DeepCopy()
	cc, err := New(completed, nil, tmpl, c.MaxPods, c.Exclude)
	if err != nil {
		out.ErrorMessage = err.Error()
		return
	}
	t0 := time.Now()
	if err := cc.SyncWithClient(client); err != nil {
		out.ErrorMessage = err.Error()
		return
	}
	out.SyncSeconds = time.Since(t0).Seconds()
	t1 := time.Now()
	if err := cc.Run(); err != nil {
		out.ErrorMessage = err.Error()
		return
	}
	out.RunSeconds = time.Since(t1).Seconds()
	rep := cc.Report()
	out.Replicas = rep.Status.Replicas
	if rep.Status.FailReason != nil {
		out.FailType = rep.Status.FailReason.FailType
		out.FailMessage = rep.Status.FailReason.FailMessage
	}
	counts := map[string]int{}
	for _, p := range cc.ScheduledPods() {
		out.Sequence = append(out.Sequence, p.Spec.NodeName)
		counts[p.Spec.NodeName]++
	}
	for n, k := range counts {
		out.PerNode = append(out.PerNode, ccrefNodeCount{NodeName: n, Replicas: k})
	}
	sort.Slice(out.PerNode, func(i, j int) bool { return out.PerNode[i].NodeName < out.PerNode[j].NodeName })
	cc.Close()
	return
}

func TestCCRefHarness(t *testing.T) {
	dir, outDir := os.Getenv("CCREF_CASES"), os.Getenv("CCREF_OUT")
	if dir == "" || outDir == "" {
		t.Skip("CCREF_CASES / CCREF_OUT not set")
	}
	files, err := filepath.Glob(filepath.Join(dir, "*.json"))
	if err != nil || len(files) == 0 {
		t.Fatalf("no case files under %s (%v)", dir, err)
	}
	sort.Strings(files)
	if err := os.MkdirAll(outDir, 0o755); err != nil {
		t.Fatal(err)
	}
	for _, f := range files {
		raw, err := os.ReadFile(f)
		if err != nil {
			t.Fatal(err)
		}
		var c ccrefCase
		if err := json.Unmarshal(raw, &c); err != nil {
			t.Fatalf("%s: %v", f, err)
		}
		if c.Name == "" {
			c.Name = strings.TrimSuffix(filepath.Base(f), ".json")
		}
		out := ccrefRun(&c)
		out.GoMaxProcs = goruntimeGOMAXPROCS()
		enc, _ := json.MarshalIndent(out, "", " ")
		if err := os.WriteFile(filepath.Join(outDir, c.Name+".out.json"), enc, 0o644); err != nil {
			t.Fatal(err)
		}
		t.Logf("%s: replicas=%d %s (%0.2fs)", c.Name, out.Replicas, out.FailType, out.RunSeconds)
		if out.ErrorMessage != "" {
			t.Errorf("%s: %s", c.Name, out.ErrorMessage)
		}
	}
}
