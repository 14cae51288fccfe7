/* _sp_fastcall: a minimal CPython binding of the C ABI (include/smallpebble_b200.h).
 *
 * The README training loop issues ~30 ABI calls per step and is HOST-bound on a B200: ctypes
 * spends ~2.5 us per call converting arguments.  This module calls the same exported
 * functions of libsmallpebble_b200.so through a vectorcall object that converts arguments
 * by a precomputed signature string (~0.3 us per call).  It contains no CUDA code and no
 * numerics -- the kernels and the ABI are unchanged; without it _abi.py falls back to ctypes.
 *
 *   fn = _sp_fastcall.bind(address, "ppplf p", name, on_error, counts_launch)
 *   fn(x_ptr, y_ptr, n, alpha, stream)      -> None, or calls on_error(name, rc)
 *
 * It also holds the free list in front of torch's caching allocator (array.py): a Block owns
 * one device buffer (a reference to the torch tensor + its address); when the LAST reference
 * to a Block goes away -- the arrays and views sharing it are all dead -- its tp_dealloc hands
 * the buffer to a per-size stack instead of freeing it, and pool_alloc() pops from there.
 * Doing this in tp_dealloc replaces a Python-level __del__ on every DeviceArray (38 per
 * training step) and a refcount heuristic with the reference count itself.
 *
 * Signature letters: 'p' pointer / integer of any width (int, None = NULL, or an object with
 * the buffer protocol such as a ctypes array or struct -> its address), 'f' float.
 * Calling convention: System V x86-64 and AAPCS64 both pass the first integer arguments and
 * the first floating-point arguments in two independent register files, in order, so one
 * trampoline type (18 integers, 4 floats) reaches every function of the ABI.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdlib.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define SP_MAX_INT_ARGS 22
#define SP_MAX_FLT_ARGS 4

typedef int (*sp_trampoline_t)(uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t,
                               uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t,
                               uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t,
                               uint64_t, uint64_t, uint64_t, uint64_t, float, float, float, float);

static long long g_launches = 0;

typedef struct {
  PyObject_HEAD
  vectorcallfunc vectorcall;
  sp_trampoline_t fn;
  char sig[SP_MAX_INT_ARGS + SP_MAX_FLT_ARGS + 1];
  int nargs;
  int counts_launch;
  PyObject* name;
  PyObject* on_error;
} FastFn;

static int to_u64(PyObject* o, uint64_t* out) {
  if (PyLong_Check(o)) {
    const unsigned long long v = PyLong_AsUnsignedLongLongMask(o);
    if (v == (unsigned long long)-1 && PyErr_Occurred()) return -1;
    *out = (uint64_t)v;
    return 0;
  }
  if (o == Py_None) {
    *out = 0;
    return 0;
  }
  if (PyIndex_Check(o)) { /* NumPy integer scalars and friends: their VALUE, not an address */
    PyObject* as_int = PyNumber_Index(o);
    if (!as_int) return -1;
    const unsigned long long v = PyLong_AsUnsignedLongLongMask(as_int);
    Py_DECREF(as_int);
    if (v == (unsigned long long)-1 && PyErr_Occurred()) return -1;
    *out = (uint64_t)v;
    return 0;
  }
  Py_buffer view; /* ctypes structs / arrays (sp_conv_geom, int64 shape vectors, pointer tables) */
  if (PyObject_GetBuffer(o, &view, PyBUF_SIMPLE) == 0) {
    *out = (uint64_t)(uintptr_t)view.buf; /* the caller keeps `o` alive across the call */
    PyBuffer_Release(&view);
    return 0;
  }
  return -1;
}

static PyObject* fastfn_call(PyObject* self_, PyObject* const* args, size_t nargsf,
                             PyObject* kwnames) {
  FastFn* self = (FastFn*)self_;
  const Py_ssize_t nargs = PyVectorcall_NARGS(nargsf);
  if (kwnames && PyTuple_GET_SIZE(kwnames) != 0) {
    PyErr_SetString(PyExc_TypeError, "ABI calls take positional arguments only");
    return NULL;
  }
  if (nargs != self->nargs) {
    PyErr_Format(PyExc_TypeError, "%U expects %d arguments, got %zd", self->name, self->nargs,
                 nargs);
    return NULL;
  }
  uint64_t iv[SP_MAX_INT_ARGS] = {0};
  float fv[SP_MAX_FLT_ARGS] = {0.f, 0.f, 0.f, 0.f};
  int ni = 0, nf = 0;
  for (Py_ssize_t i = 0; i < nargs; ++i) {
    if (self->sig[i] == 'f') {
      const double d = PyFloat_AsDouble(args[i]);
      if (d == -1.0 && PyErr_Occurred()) return NULL;
      fv[nf++] = (float)d;
    } else {
      if (to_u64(args[i], &iv[ni++]) != 0) return NULL;
    }
  }
  const int rc = self->fn(iv[0], iv[1], iv[2], iv[3], iv[4], iv[5], iv[6], iv[7], iv[8], iv[9],
                          iv[10], iv[11], iv[12], iv[13], iv[14], iv[15], iv[16], iv[17], iv[18],
                          iv[19], iv[20], iv[21], fv[0], fv[1], fv[2], fv[3]);
  if (rc != 0) {
    PyObject* r = PyObject_CallFunction(self->on_error, "Oi", self->name, rc);
    Py_XDECREF(r);
    if (!PyErr_Occurred()) PyErr_Format(PyExc_RuntimeError, "%U failed (code %d)", self->name, rc);
    return NULL;
  }
  if (self->counts_launch) ++g_launches;
  Py_RETURN_NONE;
}

static void fastfn_dealloc(FastFn* self) {
  Py_XDECREF(self->name);
  Py_XDECREF(self->on_error);
  Py_TYPE(self)->tp_free((PyObject*)self);
}

static PyObject* fastfn_get_name(FastFn* self, void* closure) {
  (void)closure;
  Py_INCREF(self->name);
  return self->name;
}

static PyGetSetDef fastfn_getset[] = {{"__name__", (getter)fastfn_get_name, NULL, NULL, NULL},
                                      {NULL, NULL, NULL, NULL, NULL}};

static PyTypeObject FastFnType = {
    PyVarObject_HEAD_INIT(NULL, 0).tp_name = "_sp_fastcall.FastFn",
    .tp_basicsize = sizeof(FastFn),
    .tp_dealloc = (destructor)fastfn_dealloc,
    .tp_vectorcall_offset = offsetof(FastFn, vectorcall),
    .tp_call = PyVectorcall_Call,
    .tp_flags = Py_TPFLAGS_DEFAULT | Py_TPFLAGS_HAVE_VECTORCALL,
    .tp_getset = fastfn_getset,
    .tp_doc = "status-checked call of one export of libsmallpebble_b200.so",
};

static PyObject* sp_bind(PyObject* mod, PyObject* args) {
  (void)mod;
  unsigned long long address;
  const char* sig;
  PyObject *name, *on_error;
  int counts_launch;
  if (!PyArg_ParseTuple(args, "KsUOp", &address, &sig, &name, &on_error, &counts_launch))
    return NULL;
  const size_t n = strlen(sig);
  size_t ni = 0, nf = 0;
  for (size_t i = 0; i < n; ++i) {
    if (sig[i] == 'f')
      ++nf;
    else if (sig[i] == 'p')
      ++ni;
    else {
      PyErr_SetString(PyExc_ValueError, "signature letters are 'p' and 'f'");
      return NULL;
    }
  }
  if (ni > SP_MAX_INT_ARGS || nf > SP_MAX_FLT_ARGS || address == 0) {
    PyErr_SetString(PyExc_ValueError, "unsupported signature or null address");
    return NULL;
  }
  FastFn* f = PyObject_New(FastFn, &FastFnType);
  if (!f) return NULL;
  f->vectorcall = fastfn_call;
  f->fn = (sp_trampoline_t)(uintptr_t)address;
  memcpy(f->sig, sig, n + 1);
  f->nargs = (int)n;
  f->counts_launch = counts_launch;
  Py_INCREF(name);
  f->name = name;
  Py_INCREF(on_error);
  f->on_error = on_error;
  return (PyObject*)f;
}

/* ---------------------------------------------------------------------------------------
 * Block pool
 * ------------------------------------------------------------------------------------- */
typedef struct {
  PyObject* tensor;
  uint64_t ptr;
} Slot;

typedef struct {
  Py_ssize_t nbytes; /* 0 = unused bucket */
  Slot* slots;
  int count, cap;
} Bucket;

#define SP_POOL_BUCKETS 512 /* distinct buffer sizes (open addressing); beyond that: no recycling */
static Bucket g_buckets[SP_POOL_BUCKETS];
static long long g_pool_bytes = 0;
static long long g_pool_max_bytes = 2LL << 30;
static int g_pool_bypass = 0;

static Bucket* bucket_for(Py_ssize_t nbytes, int create) {
  size_t h = ((size_t)nbytes * 0x9E3779B97F4A7C15ull) >> 40;
  for (int probe = 0; probe < SP_POOL_BUCKETS; ++probe) {
    Bucket* b = &g_buckets[(h + (size_t)probe) % SP_POOL_BUCKETS];
    if (b->nbytes == nbytes) return b;
    if (b->nbytes == 0) {
      if (!create) return NULL;
      b->nbytes = nbytes;
      return b;
    }
  }
  return NULL;
}

typedef struct {
  PyObject_HEAD
  PyObject* tensor;
  unsigned long long ptr;
  Py_ssize_t nbytes;
  int recycle;
} Block;

static void block_dealloc(Block* self) {
  PyObject* tensor = self->tensor;
  self->tensor = NULL;
  if (tensor && self->recycle && !g_pool_bypass &&
      g_pool_bytes + (long long)self->nbytes <= g_pool_max_bytes) {
    Bucket* b = bucket_for(self->nbytes, 1);
    if (b) {
      if (b->count == b->cap) {
        const int cap = b->cap ? 2 * b->cap : 8;
        Slot* grown = (Slot*)realloc(b->slots, (size_t)cap * sizeof(Slot));
        if (grown) {
          b->slots = grown;
          b->cap = cap;
        }
      }
      if (b->count < b->cap) {
        b->slots[b->count].tensor = tensor; /* the reference moves into the pool */
        b->slots[b->count].ptr = self->ptr;
        ++b->count;
        g_pool_bytes += (long long)self->nbytes;
        tensor = NULL;
      }
    }
  }
  Py_XDECREF(tensor);
  Py_TYPE(self)->tp_free((PyObject*)self);
}

static PyMemberDef block_members[] = {
    {"ptr", Py_T_ULONGLONG, offsetof(Block, ptr), Py_READONLY, "device address"},
    {"tensor", Py_T_OBJECT_EX, offsetof(Block, tensor), Py_READONLY, "owning torch tensor"},
    {"nbytes", Py_T_PYSSIZET, offsetof(Block, nbytes), Py_READONLY, "size in bytes"},
    {NULL, 0, 0, 0, NULL}};

static PyTypeObject BlockType = {
    PyVarObject_HEAD_INIT(NULL, 0).tp_name = "_sp_fastcall.Block",
    .tp_basicsize = sizeof(Block),
    .tp_dealloc = (destructor)block_dealloc,
    .tp_flags = Py_TPFLAGS_DEFAULT,
    .tp_members = block_members,
    .tp_doc = "one recyclable device buffer",
};

static PyObject* make_block(PyObject* tensor /* stolen */, unsigned long long ptr,
                            Py_ssize_t nbytes, int recycle) {
  Block* blk = PyObject_New(Block, &BlockType);
  if (!blk) {
    Py_DECREF(tensor);
    return NULL;
  }
  blk->tensor = tensor;
  blk->ptr = ptr;
  blk->nbytes = nbytes;
  blk->recycle = recycle;
  return (PyObject*)blk;
}

/* pool_alloc(nbytes) -> Block from the free list, or None (caller allocates + block_new) */
static PyObject* sp_pool_alloc(PyObject* mod, PyObject* arg) {
  (void)mod;
  const Py_ssize_t nbytes = PyLong_AsSsize_t(arg);
  if (nbytes == -1 && PyErr_Occurred()) return NULL;
  if (!g_pool_bypass) {
    Bucket* b = bucket_for(nbytes, 0);
    if (b && b->count > 0) {
      --b->count;
      g_pool_bytes -= (long long)nbytes;
      return make_block(b->slots[b->count].tensor, b->slots[b->count].ptr, nbytes, 1);
    }
  }
  Py_RETURN_NONE;
}

/* block_new(tensor, ptr, nbytes, recycle) -> Block */
static PyObject* sp_block_new(PyObject* mod, PyObject* args) {
  (void)mod;
  PyObject* tensor;
  unsigned long long ptr;
  Py_ssize_t nbytes;
  int recycle;
  if (!PyArg_ParseTuple(args, "OKnp", &tensor, &ptr, &nbytes, &recycle)) return NULL;
  Py_INCREF(tensor);
  return make_block(tensor, ptr, nbytes, recycle);
}

static PyObject* sp_pool_trim(PyObject* mod, PyObject* noargs) {
  (void)mod;
  (void)noargs;
  for (int i = 0; i < SP_POOL_BUCKETS; ++i) {
    Bucket* b = &g_buckets[i];
    while (b->count > 0) {
      --b->count;
      PyObject* t = b->slots[b->count].tensor;
      b->slots[b->count].tensor = NULL;
      Py_XDECREF(t);
    }
  }
  g_pool_bytes = 0;
  Py_RETURN_NONE;
}

static PyObject* sp_pool_bypass(PyObject* mod, PyObject* arg) {
  (void)mod;
  const long v = PyLong_AsLong(arg);
  if (v == -1 && PyErr_Occurred()) return NULL;
  g_pool_bypass += (int)v; /* +1 on enter, -1 on exit (nests) */
  return PyLong_FromLong(g_pool_bypass);
}

static PyObject* sp_pool_stats(PyObject* mod, PyObject* noargs) {
  (void)mod;
  (void)noargs;
  long long blocks = 0;
  for (int i = 0; i < SP_POOL_BUCKETS; ++i) blocks += g_buckets[i].count;
  return Py_BuildValue("LL", blocks, g_pool_bytes);
}

static PyObject* sp_launches(PyObject* mod, PyObject* noargs) {
  (void)mod;
  (void)noargs;
  return PyLong_FromLongLong(g_launches);
}

static PyObject* sp_add_launches(PyObject* mod, PyObject* arg) {
  (void)mod;
  const long long v = PyLong_AsLongLong(arg);
  if (v == -1 && PyErr_Occurred()) return NULL;
  g_launches += v;
  Py_RETURN_NONE;
}

static PyMethodDef methods[] = {
    {"bind", sp_bind, METH_VARARGS, "bind(address, signature, name, on_error, counts_launch)"},
    {"launches", sp_launches, METH_NOARGS, "kernel-launching ABI calls made through FastFn objects"},
    {"add_launches", sp_add_launches, METH_O, "account for launches replayed by a CUDA graph"},
    {"pool_alloc", sp_pool_alloc, METH_O, "pool_alloc(nbytes) -> recycled Block or None"},
    {"block_new", sp_block_new, METH_VARARGS, "block_new(tensor, ptr, nbytes, recycle) -> Block"},
    {"pool_trim", sp_pool_trim, METH_NOARGS, "release every pooled buffer"},
    {"pool_bypass", sp_pool_bypass, METH_O, "pool_bypass(+1 / -1): nestable bypass of the pool"},
    {"pool_stats", sp_pool_stats, METH_NOARGS, "(pooled blocks, pooled bytes)"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moduledef = {PyModuleDef_HEAD_INIT, "_sp_fastcall",
                                       "fast binding of the smallpebble_b200 C ABI", -1, methods,
                                       NULL, NULL, NULL, NULL};

PyMODINIT_FUNC PyInit__sp_fastcall(void) {
  if (PyType_Ready(&FastFnType) < 0) return NULL;
  if (PyType_Ready(&BlockType) < 0) return NULL;
  return PyModule_Create(&moduledef);
}
