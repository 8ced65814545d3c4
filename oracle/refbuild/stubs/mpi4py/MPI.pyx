# Serial stand-in implementation (see mpi4py/__init__.py in this stub tree).
import numpy as np
import copy as _copy


cdef class Status:
    def __init__(self):
        self.source = 0
        self.tag = 0
        self.error = 0

    def Get_source(self):
        return self.source

    def Get_tag(self):
        return self.tag


cdef class Datatype:
    def __init__(self, name=''):
        self.name = name


cdef class Op:
    def __init__(self, name=''):
        self.name = name


cdef class Request:
    def __init__(self, payload=None):
        self.payload = payload

    def Wait(self, status=None):
        return True

    def wait(self, status=None):
        return self.payload

    def Test(self, status=None):
        return True

    @staticmethod
    def Waitall(requests, statuses=None):
        return True

    @staticmethod
    def waitall(requests, statuses=None):
        return [r.payload for r in requests]


cdef class Prequest(Request):
    pass


cdef class Message:
    pass


cdef class Info:
    pass


cdef class Group:
    pass


cdef class Win:
    pass


cdef class File:
    pass


IN_PLACE = '__IN_PLACE__'
SUM = Op('SUM')
MAX = Op('MAX')
MIN = Op('MIN')
LOR = Op('LOR')
LAND = Op('LAND')
PROD = Op('PROD')
REAL = Datatype('REAL')
DOUBLE = Datatype('DOUBLE')
INT = Datatype('INT')
BOOL = Datatype('BOOL')
BYTE = Datatype('BYTE')
UNDEFINED = -32766
ANY_SOURCE = -1
ANY_TAG = -1
PROC_NULL = -2
ROOT = -4
LOCK_EXCLUSIVE = 1
LOCK_SHARED = 2
MODE_NOCHECK = 1
MODE_WRONLY = 4
MODE_CREATE = 1
THREAD_SINGLE = 0
THREAD_FUNNELED = 1
THREAD_SERIALIZED = 2
THREAD_MULTIPLE = 3


def _buf(b):
    if isinstance(b, (list, tuple)):
        return b[0]
    return b


cdef class Comm:
    def __init__(self, rank=0, size=1):
        self.rank = rank
        self.size = size

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.size

    def Barrier(self):
        pass

    def barrier(self):
        pass

    def Dup(self):
        return self.__class__(self.rank, self.size)

    def Split(self, color=0, key=0):
        return self.__class__(0, 1)

    def Free(self):
        pass

    def bcast(self, obj=None, root=0):
        return obj

    def Bcast(self, buf, root=0):
        pass

    def gather(self, obj, root=0):
        return [obj]

    def allgather(self, obj):
        return [obj]

    def scatter(self, objs, root=0):
        return objs[0]

    def reduce(self, obj, op=SUM, root=0):
        return obj

    def allreduce(self, obj, op=SUM):
        return obj

    def scan(self, obj, op=SUM):
        return obj

    def exscan(self, obj, op=SUM):
        return None

    def alltoall(self, objs):
        return objs

    def Allreduce(self, sendbuf, recvbuf, op=SUM):
        if isinstance(sendbuf, str) and sendbuf == IN_PLACE:
            return
        np.asarray(_buf(recvbuf))[...] = np.asarray(_buf(sendbuf))

    def Reduce(self, sendbuf, recvbuf, op=SUM, root=0):
        if isinstance(sendbuf, str) and sendbuf == IN_PLACE:
            return
        if recvbuf is not None:
            np.asarray(_buf(recvbuf))[...] = np.asarray(_buf(sendbuf))

    def Allgather(self, sendbuf, recvbuf):
        if isinstance(sendbuf, str) and sendbuf == IN_PLACE:
            return
        np.asarray(_buf(recvbuf)).flat[:] = np.asarray(_buf(sendbuf)).flat[:]

    def Gather(self, sendbuf, recvbuf, root=0):
        if recvbuf is not None:
            np.asarray(_buf(recvbuf)).flat[:] = np.asarray(_buf(sendbuf)).flat[:]

    def Gatherv(self, sendbuf, recvbuf, root=0):
        if recvbuf is not None:
            r = np.asarray(_buf(recvbuf))
            s = np.asarray(_buf(sendbuf))
            r.flat[:s.size] = s.flat[:]

    def Allgatherv(self, sendbuf, recvbuf):
        r = np.asarray(_buf(recvbuf))
        s = np.asarray(_buf(sendbuf))
        r.flat[:s.size] = s.flat[:]

    def Alltoall(self, sendbuf, recvbuf):
        np.asarray(_buf(recvbuf)).flat[:] = np.asarray(_buf(sendbuf)).flat[:]

    def Alltoallv(self, sendbuf, recvbuf):
        r = np.asarray(_buf(recvbuf))
        s = np.asarray(_buf(sendbuf))
        r.flat[:s.size] = s.flat[:]

    def isend(self, obj, dest=0, tag=0):
        return Request(obj)

    def Isend(self, buf, dest=0, tag=0):
        return Request(buf)

    def irecv(self, buf=None, source=0, tag=0):
        return Request(None)

    def Irecv(self, buf, source=0, tag=0):
        return Request(None)

    def send(self, obj, dest=0, tag=0):
        raise RuntimeError('serial MPI stub: send without matching recv')

    def recv(self, buf=None, source=0, tag=0, status=None):
        raise RuntimeError('serial MPI stub: recv without matching send')

    def Abort(self, errorcode=0):
        raise SystemExit(errorcode)

    def Get_name(self):
        return 'fake'

    def Is_inter(self):
        return False


cdef class Intracomm(Comm):
    pass


cdef class Intercomm(Comm):
    pass


def fakeComm(rank, size):
    """Communicator that *claims* (rank, size) but performs identity
    collectives; partial results are summed by the caller."""
    return Intracomm(rank, size)


COMM_WORLD = Intracomm(0, 1)
COMM_SELF = Intracomm(0, 1)
COMM_NULL = None


def Is_initialized():
    return True


def Is_finalized():
    return False


def Init():
    pass


def Finalize():
    pass


def Query_thread():
    return THREAD_MULTIPLE


def Is_thread_main():
    return True


def Get_version():
    return (4, 0)


def Get_library_version():
    return 'serial stub'


def Get_processor_name():
    return 'localhost'


def get_vendor():
    return ('stub', (0, 0, 0))


def Wtime():
    import time
    return time.time()
