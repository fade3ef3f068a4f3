"""Single-rank MPI stand-in (see package docstring)."""
SUM = "sum"
MAX = "max"
DOUBLE = "double"


class _Request(object):
    def Start(self):
        pass

    def Wait(self):
        pass


class Prequest(object):
    @staticmethod
    def Startall(reqs):
        pass

    @staticmethod
    def Waitall(reqs):
        pass


class _Comm(object):
    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def allreduce(self, value, op=SUM):
        return value

    def allgather(self, value):
        return [value]

    def bcast(self, value, root=0):
        return value

    def Barrier(self):
        pass

    def Split(self, color=0, key=0):
        return self

    def Send_init(self, buf, dest, tag=0):
        return _Request()

    def Recv_init(self, buf, source, tag=0):
        return _Request()

    Ssend_init = Send_init


COMM_WORLD = _Comm()
