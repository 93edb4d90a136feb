// bnnp_host.cpp -- host-side helper of the Python sampler classes (NOT part of the C ABI in include/bnnp.h).
//
// The reference loops over the parameters in Python at every step (mcmc/sgld.py:94-105: `p.grad is None`,
// `p.grad` ...).  The B200 sampler has to look at every parameter too -- has autograd handed a gradient
// over, where does it lie, is the parameter still the view of the flat array -- and in Python that costs
// two attribute / method calls per tensor: 13 of the 21-27 us a step takes on the host for the 65 tensors of
// `googleresnet` (profiles/r02_notes.md).  This extension does the same scan over the ATen objects directly
// (no Python objects are created): ~1 us.  Everything that is not "nothing changed" goes back to the Python
// slow path (bnn_priors_b200/mcmc/_flat.py), which stays the specification; without this module the samplers
// run on the Python scan alone.
#include <torch/extension.h>

#include <vector>

namespace {

struct GradScanner {
    std::vector<at::Tensor> params;
    std::vector<int64_t> table;        // gradient address per tensor the device table holds
    std::vector<int64_t> p_ptrs;       // address of every parameter's view of the flat P array
    std::vector<at::Tensor> held;      // the gradients SUM_GG describes (strong references, like _held_grads)
    std::vector<int64_t> held_versions;

    GradScanner(std::vector<at::Tensor> ps, std::vector<int64_t> pp) : params(std::move(ps)), p_ptrs(std::move(pp)) {
        table.assign(params.size(), 0);
    }

    void set_table(const std::vector<int64_t>& t) { table = t; }

    // 0: every parameter has a gradient at the address the device table holds and is still the flat view;
    // 1: a gradient is missing or lies elsewhere (Python slow path); 2: only a parameter's storage changed.
    int scan() const {
        const size_t n = params.size();
        for (size_t i = 0; i < n; ++i) {
            const at::Tensor& g = params[i].grad();
            // (layout and type again, not only the address: a transposed view of a square gradient starts at the
            //  same byte -- the Python slow path then copies it)
            if (!g.defined() || reinterpret_cast<int64_t>(g.data_ptr()) != table[i] || !g.is_contiguous() ||
                g.scalar_type() != at::kFloat)
                return 1;
        }
        for (size_t i = 0; i < n; ++i)
            if (reinterpret_cast<int64_t>(params[i].data_ptr()) != p_ptrs[i]) return 2;
        return 0;
    }

    // remember which gradient tensors (and which versions of them) the sums of the last launch describe
    bool capture() {
        const size_t n = params.size();
        held.clear();
        held_versions.clear();
        held.reserve(n);
        held_versions.reserve(n);
        for (size_t i = 0; i < n; ++i) {
            const at::Tensor& g = params[i].grad();
            if (!g.defined()) {
                held.clear();
                held_versions.clear();
                return false;
            }
            held.push_back(g);
            held_versions.push_back(static_cast<int64_t>(g._version()));
        }
        return true;
    }

    void drop() {
        held.clear();
        held_versions.clear();
    }

    bool fresh() const {
        const size_t n = params.size();
        if (held.size() != n) return false;
        for (size_t i = 0; i < n; ++i) {
            const at::Tensor& g = params[i].grad();
            if (!g.defined() || g.unsafeGetTensorImpl() != held[i].unsafeGetTensorImpl() ||
                static_cast<int64_t>(g._version()) != held_versions[i])
                return false;
        }
        return true;
    }

    // sum of the parameters' version counters: changes whenever somebody writes a parameter in place
    int64_t params_version() const {
        int64_t v = 0;
        for (const at::Tensor& p : params) v += static_cast<int64_t>(p._version());
        return v;
    }

    // optimizer.zero_grad(set_to_none=True): p.grad = None for every parameter except `keep` (fused
    // hyper-parameters keep their zeroed flat view)
    void drop_grads(const std::vector<int64_t>& keep) {
        const size_t n = params.size();
        size_t k = 0;
        for (size_t i = 0; i < n; ++i) {
            if (k < keep.size() && keep[k] == static_cast<int64_t>(i)) {
                ++k;
                continue;
            }
            params[i].mutable_grad().reset();
        }
        drop();
    }
};

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    py::class_<GradScanner>(m, "GradScanner")
        .def(py::init<std::vector<at::Tensor>, std::vector<int64_t>>())
        .def("set_table", &GradScanner::set_table)
        .def("scan", &GradScanner::scan)
        .def("capture", &GradScanner::capture)
        .def("drop", &GradScanner::drop)
        .def("fresh", &GradScanner::fresh)
        .def("params_version", &GradScanner::params_version)
        .def("drop_grads", &GradScanner::drop_grads);
}
