"""NumPy restatement of the reference's prioritized replay memory -- TEST INFRASTRUCTURE ONLY.

Follows distributed_queue/buffer_queue.py:326-369 (``SumTree``) and :371-415 (``Memory``): array-embedded binary sum
tree of 2*capacity-1 float64 nodes (leaves last), priority (error + 0.001) ** 0.6, stratified sampling of n segments
of total/n, importance weights (n_entries * p/total) ** -beta normalised by their maximum, beta annealed 0.4 -> 1 by
0.001 per ``sample`` call.  The uniform draw of ``random.uniform(a, b)`` is taken as a + (b - a) * u with u supplied
by the caller, which is what CPython's ``random.uniform`` computes from ``random.random()``.

PINNED: the reference's ``SumTree`` / ``Memory`` are plain NumPy, so ``tests/test_oracle_refexec.py`` executes them as
they are (CPython ``random`` seeded) and this restatement equals them bit for bit (tree nodes, sampled indices, weights,
beta); further pinned by hand-computed totals and the invariants (every internal node = sum of its children, the
sampled leaf contains the drawn mass).
"""
import numpy as np


class SumTreeNP:
    def __init__(self, capacity):
        self.capacity = int(capacity)
        self.nodes = np.zeros(2 * self.capacity - 1, dtype=np.float64)
        self.cursor = 0
        self.n_entries = 0

    def set_leaf(self, node, p):
        """buffer_queue.py:361-364 + :334-338: write the leaf, add the difference to every ancestor (leaf -> root)."""
        delta = p - self.nodes[node]
        self.nodes[node] = p
        while node != 0:
            node = (node - 1) // 2
            self.nodes[node] += delta

    def add(self, p):
        """buffer_queue.py:351-359 -> data index used."""
        slot = self.cursor
        self.set_leaf(slot + self.capacity - 1, p)
        self.cursor = (slot + 1) % self.capacity if slot + 1 >= self.capacity else slot + 1
        self.n_entries = min(self.n_entries + 1, self.capacity)
        return slot

    def find(self, mass):
        """buffer_queue.py:340-347,366-369: descend left when mass <= left subtree sum, else right with the rest."""
        node = 0
        while 2 * node + 1 < len(self.nodes):
            left = 2 * node + 1
            if mass <= self.nodes[left]:
                node = left
            else:
                mass = mass - self.nodes[left]
                node = left + 1
        return node

    def total(self):
        return self.nodes[0]


class MemoryNP:
    e, a, beta0, beta_step = 0.001, 0.6, 0.4, 0.001        # buffer_queue.py:372-375

    def __init__(self, capacity):
        self.tree = SumTreeNP(capacity)
        self.beta = self.beta0

    def priority(self, error):
        return (np.float64(error) + self.e) ** self.a       # :383-384

    def add(self, error):
        return self.tree.add(self.priority(error))

    def sample(self, n, u01):
        """:390-411 -> (tree indices, data indices, priorities, is_weight)."""
        seg = self.tree.total() / n
        self.beta = np.min([1.0, self.beta + self.beta_step])
        idxs, prios = [], []
        for i in range(n):
            lo, hi = seg * i, seg * (i + 1)
            node = self.tree.find(lo + (hi - lo) * u01[i])
            idxs.append(node)
            prios.append(self.tree.nodes[node])
        probs = np.asarray(prios) / self.tree.total()
        w = np.power(self.tree.n_entries * probs, -self.beta)
        w /= w.max()
        idxs = np.asarray(idxs, np.int64)
        return idxs, idxs - self.tree.capacity + 1, np.asarray(prios), w

    def update(self, node, error):
        self.tree.set_leaf(int(node), self.priority(error))  # :413-415
