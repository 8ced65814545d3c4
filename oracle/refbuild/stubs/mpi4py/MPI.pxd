# Serial stand-in declarations (see mpi4py/__init__.py in this stub tree).
cdef class Status:
    cdef public int source, tag, error

cdef class Datatype:
    cdef public object name

cdef class Op:
    cdef public object name

cdef class Request:
    cdef public object payload

cdef class Prequest(Request):
    pass

cdef class Message:
    pass

cdef class Info:
    pass

cdef class Group:
    pass

cdef class Win:
    cdef public object mem

cdef class File:
    pass

cdef class Comm:
    cdef public int rank, size

cdef class Intracomm(Comm):
    pass

cdef class Intercomm(Comm):
    pass
