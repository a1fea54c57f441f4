// Host-side glue of the Python drop-in layer.  Not part of the C ABI (include/b200det.h) and no device code: it is the
// tail of `postprocess.non_max_suppression` — the part that runs on the host AFTER the counts event, while the GPU has only the
// emit kernel (27 us at the headline) left to hide it behind.  In Python that tail is 64 x `Tensor.resize_` = 28 us of argument
// parsing and dispatch (tools/api_profile.py); here it is one call.
//
// finish_views(views, counts_ptr, cols): `views` are B views [n_pad, cols] (cols == 0: [n_pad]) of the padded result the
// emit kernel is writing, made while the GPU was busy; `counts_ptr` is the address of the pinned host buffer the NMS stage wrote
// the B int32 counts into.  Every view is shrunk in place to its count (same storage, same offset: no allocation, no copy);
// images without detections become None, as the YOLO reference returns them (model/YOLOV3.py:306,333), or stay as empty
// [0, cols] tensors with keep_empty (the SSD / RetinaNet reference stacks an empty result, model/SSD.py:304-310).
#include <torch/extension.h>

#include <cstdint>
#include <vector>

static py::object finish_views(const py::list& views, const int64_t counts_ptr, const int64_t cols, const bool keep_empty) {
    const int32_t* counts = reinterpret_cast<const int32_t*>(static_cast<intptr_t>(counts_ptr));
    const Py_ssize_t n = PyList_GET_SIZE(views.ptr());
    PyObject* out = PyList_New(n);
    if (!out) throw py::error_already_set();
    for (Py_ssize_t b = 0; b < n; ++b) {
        const int64_t k = counts[b];
        PyObject* o = Py_None;
        if (k > 0 || (keep_empty && k == 0)) {
            o = PyList_GET_ITEM(views.ptr(), b);                       // borrowed; the result list shares the very same objects
            if (!THPVariable_Check(o)) {
                Py_DECREF(out);
                throw py::type_error("finish_views: `views` must be a list of tensors");
            }
            const at::Tensor& v = THPVariable_Unpack(o);
            if (k > v.size(0) || !v.is_contiguous() || v.dim() != (cols > 0 ? 2 : 1)) {
                Py_DECREF(out);
                TORCH_CHECK(false, "finish_views: image ", (int64_t)b, ": count ", k, " does not fit its padded view");
            }
            // same storage, same offset, fewer rows: only the size changes (no dispatch, no allocation)
            if (cols > 0) v.unsafeGetTensorImpl()->set_sizes_contiguous({k, cols});
            else v.unsafeGetTensorImpl()->set_sizes_contiguous({k});
        }
        Py_INCREF(o);
        PyList_SET_ITEM(out, b, o);
    }
    return py::reinterpret_steal<py::object>(out);
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("finish_views", &finish_views, "shrink the per-image views of a padded NMS result to their counts", py::arg("views"),
          py::arg("counts_ptr"), py::arg("cols"), py::arg("keep_empty") = false);
}
