"""Cluster tree and admissible cluster pairs of the H2 format (host side).

The structure decides which entries the near field holds and which far-field kernel blocks
(`nonlocalBuilder.getFarFieldBlocks`, pnb_farfield_blocks) exist, so it has to be the reference's, node for node:

* per-DoF boxes and mesh sizes            getDoFBoxesAndCells (clusterMethodCy.pyx:3922-3977), getHVector
                                          (nonlocalAssembly_{SCALAR}.pxi:2386-2399)
* parameters                              getH2RefinementParams (nonlocalAssembly_{SCALAR}.pxi:2983-3046)
* node boxes / interpolation orders       tree_node.init (clusterMethodCy.pyx:177-206)
* median bisection, node ids              tree_node.refine (clusterMethodCy.pyx:354-663)
* admissibility, lazy refinement          getAdmissibleClusters (clusterMethodCy.pyx:4044-4136), box distance
                                          (interactionDomains.pyx:311-324), diamBox (clusterMethodCy.pyx:121-127)
* trimming                                trimTree / tree_node.trim (clusterMethodCy.pyx:4218-4240, 1461-1487)

Serial, infinite horizon, constant kernel parameters, MEDIAN refinement (the defaults).  Everything here is index and
comparison work on a few thousand boxes; it stays on the host (numpy).
"""
from math import ceil, log, sqrt

import numpy as np


class refinementParams:
    def __init__(self, mesh, kernel, target_order, params={}):
        singularity = kernel.singularityValue
        self.targetOrder = target_order
        self.meshDiam = mesh.diam
        self.eta = params.get('eta', 3.)
        self.maxSingularity = singularity
        loggamma = abs(log(0.25))

        def order(h):
            return max(ceil((2*target_order+max(-singularity, 2))*abs(log(h/mesh.diam))/loggamma/3.), 2)
        iO = params.get('interpolation_order', None)
        self.interpolation_order = order(mesh.hmin) if iO is None else iO
        mL = params.get('maxLevels', None)
        self.maxLevels = 200 if mL is None else mL
        mFFBS = params.get('minFarFieldBlockSize', None)
        self.farFieldInteractionSize = -1 if mFFBS is None else mFFBS
        mCS = params.get('minClusterSize', None)
        if mCS is None:
            io = order(mesh.h) if self.farFieldInteractionSize < 0 else self.interpolation_order
            self.minSize = int(io**mesh.dim//2)
        else:
            self.minSize = mCS
        if params.get('refinementType', 'MEDIAN') not in ('MEDIAN', 'median'):
            raise NotImplementedError('only MEDIAN refinement')
        if params.get('splitEveryDim', False):
            raise NotImplementedError('splitEveryDim')
        self.attemptRefinement = True


def dof_boxes(mesh, dm):
    """bounding box of the cells around every DoF, [num_dofs, dim, 2]"""
    v = mesh.vertices[mesh.cells]                     # cells x (dim+1) x dim
    lo, hi = v.min(axis=1), v.max(axis=1)
    boxes = np.empty((dm.num_dofs, mesh.dim, 2))
    boxes[:, :, 0] = np.inf
    boxes[:, :, 1] = -np.inf
    for k in range(dm.dofs.shape[1]):
        m = dm.dofs[:, k] >= 0
        np.minimum.at(boxes[:, :, 0], dm.dofs[m, k], lo[m])
        np.maximum.at(boxes[:, :, 1], dm.dofs[m, k], hi[m])
    return boxes


def dof_h(mesh, dm):
    """smallest cell size in the patch of every DoF"""
    h = np.full(dm.num_dofs, np.inf)
    for k in range(dm.dofs.shape[1]):
        m = dm.dofs[:, k] >= 0
        np.minimum.at(h, dm.dofs[m, k], mesh.hVector[m])
    return h


def dist_boxes(b1, b2):
    d = 0.
    for i in range(b1.shape[0]):
        if b1[i, 0] > b2[i, 0]:
            gap = b1[i, 0]-b2[i, 1]
        else:
            gap = b2[i, 0]-b1[i, 1]
        d += max(gap, 0)**2
    return sqrt(d)


def diam_box(b):
    d = 0.
    for i in range(b.shape[0]):
        d += (b[i, 1]-b[i, 0])**2
    return sqrt(d)


class dofArray(np.ndarray):
    """sorted dof indices of a cluster with the `toSet()` of the reference's indexSet (tests/test_fracLapl.py:174)"""

    def toSet(self):
        return set(self.tolist())


class tree_node:
    def __init__(self, parent, dofs, data, mixed_node=False):
        self.parent = parent
        self.children = []
        self._dofs = dofs                 # sorted int array (leaves)
        self.data = data                  # (boxes, coords, hVector, refParams)
        self.mixed_node = mixed_node
        self.canBeAssembled = True
        self.id = 0
        self.levelNo = 0 if parent is None else parent.levelNo+1
        boxes, coords, hVector, rp = data
        self.dim = boxes.shape[1]
        self.box = np.empty((self.dim, 2))
        self.box[:, 0] = boxes[dofs, :, 0].min(axis=0)
        self.box[:, 1] = boxes[dofs, :, 1].max(axis=0)
        self.hmin = hVector[dofs].min()
        if rp.farFieldInteractionSize < 0:
            self.interpolation_order = int(max(ceil((2*rp.targetOrder+max(-rp.maxSingularity, 2))
                                                    * abs(log(self.hmin/rp.meshDiam))/abs(log(0.25))/3.), 2))
        else:
            self.interpolation_order = rp.interpolation_order
        self._num_dofs = dofs.shape[0]
        self.irregularLevelsOffset = 0
        self._unsplittable = False

    isLeaf = property(lambda self: len(self.children) == 0)
    num_dofs = property(lambda self: self._num_dofs)

    @property
    def dofs(self):
        if self.isLeaf:
            return self._dofs.view(dofArray)
        return np.unique(np.concatenate([c.dofs for c in self.children])).view(dofArray)

    def root(self):
        n = self
        while n.parent is not None:
            n = n.parent
        return n

    def _num_root_children(self):
        root = self.root()

        def count(n, off):
            if off > 1:
                return sum(count(c, off-1) for c in n.children)
            return len(n.children)
        return count(root, root.irregularLevelsOffset)

    def get_tree_nodes(self):
        yield self
        for c in self.children:
            yield from c.get_tree_nodes()

    # ---- the passes of the H2 product, node by node (tree_node.upwardPass / downwardPass / resetCoefficientsDown,
    # clusterMethodCy.pyx:1093-1176), for callers that drive them by hand like tests/test_fracLapl.py:176-186 of the
    # reference.  Device tensors throughout (the cluster bases are attached by H2Matrix); H2Matrix.matvec itself runs
    # the batched kernels of csrc/pnb_h2.cuh.
    def _dev(self, array, key):
        import torch
        cache = self.__dict__.setdefault('_tensors', {})
        if key not in cache:
            cache[key] = torch.as_tensor(np.ascontiguousarray(array), device=self.root()._device)
        return cache[key]

    def upwardPass_py(self, x, componentNo=0, skip_leaves=False):
        import torch
        if not isinstance(x, torch.Tensor):
            x = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64), device=self.root()._device)
        if self.isLeaf:
            if not skip_leaves:
                self.coefficientsUp = self._dev(self.value, 'V').t().mv(x[self._dev(self.dofs, 'dofs')])
        else:
            acc = None
            for c in self.children:
                c.upwardPass_py(x, componentNo, skip_leaves)
                t = c._dev(c.transferOperator, 'T').mv(c.coefficientsUp)
                acc = t if acc is None else acc+t
            self.coefficientsUp = acc

    def resetCoefficientsDown_py(self, vecValued=False):
        import torch
        for n in self.get_tree_nodes():
            n.coefficientsDown = torch.zeros(n.interpolation_order**n.dim, dtype=torch.float64, device=self.root()._device)

    def downwardPass_py(self, y, componentNo=0):
        """adds the far-field part to y (host array: in place through a device copy; device tensor: in place)"""
        import torch
        host = not isinstance(y, torch.Tensor)
        yt = torch.as_tensor(np.ascontiguousarray(y, dtype=np.float64), device=self.root()._device) if host else y
        self._down(yt)
        if host:
            y[:] = yt.cpu().numpy()

    def _down(self, yt):
        if self.isLeaf:
            yt[self._dev(self.dofs, 'dofs')] += self._dev(self.value, 'V').mv(self.coefficientsDown)
        else:
            for c in self.children:
                c.coefficientsDown += c._dev(c.transferOperator, 'T').t().mv(self.coefficientsDown)
                c._down(yt)

    def get_tree_nodes_up_to_level(self, level):
        yield self
        if level > 0:
            for c in self.children:
                yield from c.get_tree_nodes_up_to_level(level-1)

    def leaves(self):
        if self.isLeaf:
            yield self
        for c in self.children:
            yield from c.leaves()

    def get_max_id(self):
        return max([self.id]+[c.get_max_id() for c in self.children])

    # ---- median bisection -----------------------------------------------------------------------------------
    def refine(self, recursive=True):
        if self._unsplittable:
            return
        boxes, coords, hVector, rp = self.data
        dofs = self._dofs
        n0 = dofs.shape[0]
        limit_levels, limit_size = rp.maxLevels, rp.minSize
        if (self.levelNo+1 >= limit_levels) or (n0 <= limit_size):
            self._unsplittable = True
            return
        dim = self.dim
        if dim == 1:
            m0 = np.median(coords[dofs, 0])
            first = (self.box[0, 0] <= coords[dofs, 0]) & (coords[dofs, 0] < m0)
        else:
            split = 0
            size = self.box[0, 1]-self.box[0, 0]
            for i in range(1, dim):
                if self.box[i, 1]-self.box[i, 0] > size:
                    split = i
                    size = self.box[i, 1]-self.box[i, 0]
            median = np.median(coords[dofs, split])
            sub = np.empty((dim, 2))
            sub[:, 0] = self.box[:, 0]-1e-12
            sub[:, 1] = self.box[:, 1]+1e-12
            sub[split, 1] = median
            first = np.ones(n0, dtype=bool)
            for i in range(dim):
                first &= (sub[i, 0] <= coords[dofs, i]) & (coords[dofs, i] < sub[i, 1])
        children = []
        lvl = self.levelNo
        for k, sel in enumerate((first, ~first)):
            d = dofs[sel]
            if not (d.shape[0] >= rp.minSize and d.shape[0] < n0):
                self._unsplittable = True         # the admissibility search asks again for every partner cluster
                return
            c = tree_node(self, d, self.data, mixed_node=self.mixed_node)
            if lvl > 0:
                nrc = self._num_root_children()
                lvlID = self.id-(nrc*(2**(lvl-1)-1)//(2-1)+1)
                c.id = nrc*(2**lvl-1)//(2-1)+1+2*lvlID+k
            else:
                c.id = k+1
            children.append(c)
        self.children = children
        self._dofs = None
        if recursive:
            for c in self.children:
                c.refine(recursive)

    # ---- trimming ---------------------------------------------------------------------------------------------
    def trim(self, keep):
        delNode = self.id not in keep
        newChildren = []
        delAll = True
        for c in self.children:
            cdel = c.trim(keep)
            if not cdel:
                delNode = False
                newChildren.append(c)
            delAll &= cdel
        if not self.isLeaf and len(newChildren) == 0:
            self._dofs = self.dofs
            self.children = []
        elif len(self.children) > 0 and delAll:
            adopted = []
            for c in self.children:
                for c2 in c.children:
                    adopted.append(c2)
                    c2.parent = self
            self.children = adopted
        return delNode


def build_tree(mesh, dm, kernel, target_order, params={}):
    """root cluster over all DoFs (nonlocalBuilder.getTree, serial branch, nonlocalAssembly_{SCALAR}.pxi:2541-2664)"""
    rp = refinementParams(mesh, kernel, target_order, params)
    data = (dof_boxes(mesh, dm), dm.getDoFCoordinates(), dof_h(mesh, dm), rp)
    root = tree_node(None, np.arange(dm.num_dofs), data)
    root.irregularLevelsOffset = 1
    return root


def _admissible(n1, n2, dim, Pfar, Pnear, level):
    rp = n1.data[3]
    dist = dist_boxes(n1.box, n2.box)
    diam1, diam2 = diam_box(n1.box), diam_box(n2.box)
    size = (n1.interpolation_order*n2.interpolation_order)**dim
    seems = (rp.eta*dist >= max(diam1, diam2) and not n1.mixed_node and not n2.mixed_node
             and size <= n1.num_dofs*n2.num_dofs and n1.canBeAssembled and n2.canBeAssembled)
    nnear = len(Pnear)
    added = False
    if seems:
        Pfar.setdefault(level, []).append((n1, n2))
        return True
    if rp.attemptRefinement:
        if n1.isLeaf:
            n1.refine(False)
        if n2.isLeaf:
            n2.refine(False)
    if (n1.isLeaf and n2.isLeaf) or level == rp.maxLevels:
        Pnear.append((n1, n2))
        return False
    elif size > n1.num_dofs*n2.num_dofs:
        Pnear.append((n1, n2))
        return False
    elif n1.isLeaf:
        for t2 in n2.children:
            added |= _admissible(n1, t2, dim, Pfar, Pnear, level+1)
    elif n2.isLeaf:
        for t1 in n1.children:
            added |= _admissible(t1, n2, dim, Pfar, Pnear, level+1)
    else:
        for t1 in n1.children:
            for t2 in n2.children:
                added |= _admissible(t1, t2, dim, Pfar, Pnear, level+1)
    if not added:
        # nothing below is admissible: one near-field pair for the whole block
        del Pnear[nnear:]
        Pnear.append((n1, n2))
    return added


def admissible_clusters(root, trim=True):
    """(Pnear, Pfar): near-field cluster pairs in the reference's order and far-field pairs per level
    (nonlocalBuilder.getAdmissibleClusters, serial branch, nonlocalAssembly_{SCALAR}.pxi:2842-2847)"""
    Pnear, Pfar = [], {}
    _admissible(root, root, root.dim, Pfar, Pnear, 0)
    if trim:
        keep = set()
        for a, b in Pnear:
            keep.update((a.id, b.id))
        for lvl in Pfar:
            for a, b in Pfar[lvl]:
                keep.update((a.id, b.id))
        for n in root.get_tree_nodes_up_to_level(root.irregularLevelsOffset):
            keep.add(n.id)
        root.trim(keep)
    return Pnear, Pfar
