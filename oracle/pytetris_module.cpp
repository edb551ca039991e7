/*
 * pytetris_module.cpp — TEST INFRASTRUCTURE ONLY.  Builds oracle/_ref/pyTetris*.so: the Python face of the CPU
 * oracle env, shaped like the absent hrpan/pyTetris module so the reference's UNMODIFIED agents
 * (agents/agent.py, ValueSimLP.py, Vanilla.py, cppmodule/agent.cpp) can run on it when golden vectors are
 * generated (tests/golden/gen_golden.py) and when bench.py times the reference CPU arm.  The buffer protocol
 * exposes the C++ object itself (agent.cpp:211-214,267-270,275-276 reinterpret info.ptr as Tetris*).
 */
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>
#include <pybind11/numpy.h>
#include "pyTetris/pyTetris.h"
namespace py = pybind11;

PYBIND11_MODULE(pyTetris, m) {
    py::class_<Tetris>(m, "Tetris", py::buffer_protocol())
        .def(py::init<>())
        .def(py::init<std::pair<int, int>, int, int, int>(), py::arg("boardsize") = std::pair<int, int>(20, 10),
             py::arg("actions_per_drop") = 1, py::arg("scoring") = 0, py::arg("randomizer") = 0)
        .def_buffer([](Tetris &t) {
            return py::buffer_info(reinterpret_cast<unsigned char *>(&t), 1, py::format_descriptor<unsigned char>::format(), 1,
                                   {sizeof(Tetris)}, {1});
        })
        .def("play", &Tetris::play)
        .def("reset", &Tetris::reset)
        .def("seed", &Tetris::seed)
        .def("copy_from", &Tetris::copy_from)
        .def("clone", &Tetris::clone)
        .def("equiv", &Tetris::equiv)
        .def("getState", &Tetris::getState)
        .def("_getState", &Tetris::_getState)
        .def("getScore", &Tetris::getScore)
        .def("printState", &Tetris::printState)
        .def("hash", &Tetris::hash)
        .def("get_record", &Tetris::get_record)
        .def("set_record", &Tetris::set_record)
        .def("__hash__", [](const Tetris &t) { return (py::ssize_t)(t.hash() >> 1); })
        .def("__eq__", [](const Tetris &a, const Tetris &b) { return a == b; })
        .def_readonly("end", &Tetris::end)
        .def_readonly("score", &Tetris::score)
        .def_readonly("combo", &Tetris::combo)
        .def_readonly("line_clears", &Tetris::line_clears)
        .def_property_readonly("line_stats", [](const Tetris &t) {
            py::array_t<int> a(4);
            for (int i = 0; i < 4; ++i) a.mutable_at(i) = t.line_stats[i];
            return a;
        });
}
