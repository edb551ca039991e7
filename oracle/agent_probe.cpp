// agent_probe.cpp — TEST INFRASTRUCTURE ONLY (never shipped, never on the product path).
//
// Compiles the reference's agents/cppmodule/agent.cpp UNCHANGED — by inclusion, from where it lies under /root/reference (oracle/Makefile passes
// -I.../agents/cppmodule) — and adds a second pybind11 module that scripts the INPUTS of OnlineMCTSAgent::remove_nodes (agent.cpp:619-708):
// all members of the reference's classes are public, so the probe writes the observation arrays a collection would free (visit_obs, value_obs,
// variance_obs, end_obs, state_obs) and the episode counter, and then calls the reference's own remove_nodes().  That runs, untouched:
// store_nodes (agent.cpp:777-819, incl. the policy-0 random drop), the four accumulation policies (:635-695), weighted_trimming (:710-749),
// random_trimming (:751-775) and the train callback (:698).  Why a probe: driven through a real search the memory also depends on a defect of
// TreeAgent::update_available (agent.cpp:300-301: `occupied.resize(n); occupied.insert(begin, ...)` keeps the first n STALE entries of the
// allocation history, so observations of long-dead nodes stay "occupied" in an order that depends on std::unordered_set iteration) — not
// something a re-implementation can or should reproduce.  With an empty tree (root = 0) update_available() frees every observation index
// 1..max_nodes-1, in ascending order, which is exactly the list a collection hands to store_nodes.
#include <pybind11/pybind11.h>
#include <pybind11/numpy.h>
#include <pybind11/stl.h>
#undef PYBIND11_MODULE
#define PYBIND11_MODULE(name, variable) static void reference_module_##name(pybind11::module_ &variable)
#include <agent.cpp>          // the reference's file, unmodified
#undef PYBIND11_MODULE

namespace probe {
struct Probe {
    OnlineMCTSAgent agent;
    Probe(int policy, int memory_size, int episodes_per_train, int growth, int min_visit, int max_nodes, py::function train)
        : agent(make(policy, memory_size, episodes_per_train, growth, min_visit, max_nodes, train)) { agent.root = 0; }
    static OnlineMCTSAgent make(int policy, int memory_size, int ept, int growth, int min_visit, int max_nodes, py::function train) {
        int sims = 1; bool online = true, projection = true, benchmark = false, lp = true; double gamma = 0.999; int etype = 0;
        py::function eval = train;
        return OnlineMCTSAgent(sims, max_nodes, online, policy, memory_size, ept, growth, min_visit, projection, gamma, benchmark, eval, etype, train, lp);
    }
    // one collection: the observations idx[] (ascending, 1 <= idx < max_nodes) are what it frees
    void collect(std::vector<int> idx, std::vector<int> visit, std::vector<float> value, std::vector<float> variance, std::vector<int> end,
                 py::array_t<int8_t, py::array::c_style | py::array::forcecast> states, int current_episode) {
        auto s = states.unchecked<2>();
        for (size_t k = 0; k < idx.size(); ++k) {
            const int o = idx[k];
            agent.visit_obs[o] = visit[k]; agent.value_obs[o] = value[k]; agent.variance_obs[o] = variance[k]; agent.end_obs[o] = end[k] != 0;
            agent.state_obs[o].resize(200);                       // _new_node assigns the 200-cell state here (agent.cpp:247)
            for (int c = 0; c < 200; ++c) agent.state_obs[o][c] = (char)s(k, c);
        }
        agent.current_episode = current_episode;
        agent.remove_nodes();
    }
    int memory_index() const { return agent.memory_index; }
    int n_trains() const { return agent.n_trains; }
    double drop_prob() const { return agent.memory_drop_prob; }
};
}  // namespace probe

extern "C" PYBIND11_EXPORT PyObject *PyInit_agent_probe();
static void init_probe(py::module_ &m) {
    py::class_<probe::Probe>(m, "Probe")
        .def(py::init<int, int, int, int, int, int, py::function>())
        .def("collect", &probe::Probe::collect)
        .def("memory_index", &probe::Probe::memory_index)
        .def("n_trains", &probe::Probe::n_trains)
        .def("drop_prob", &probe::Probe::drop_prob);
}
static py::module_::module_def probe_def;
extern "C" PYBIND11_EXPORT PyObject *PyInit_agent_probe() {
    PYBIND11_CHECK_PYTHON_VERSION
    PYBIND11_ENSURE_INTERNALS_READY
    auto m = py::module_::create_extension_module("agent_probe", nullptr, &probe_def);
    try { init_probe(m); return m.ptr(); }
    PYBIND11_CATCH_INIT_EXCEPTIONS
}
