"""CPU: where the variable order `dense_0..dense_9, rgb, sigma` comes from.

Checkpoints (`core/ops.py:146-149`, `core/model.py:259-276`) and `optimizer.variables()` depend on the order of
`model.trainable_variables` of the reference's functional model (`core/model.py:334-394`). That order is NOT the
creation order of the layers (sigma is created before dense_8, dense_9 and rgb): a Keras functional model lists
its layers by DECREASING DEPTH from the outputs, ties broken by the order in which a depth-first walk from the outputs
first meets them (keras/engine/functional.py, `_map_graph_network` / `_build_map`). Keras is not installable here, so
the algorithm is restated below from its published source and run on the reference's own graph; the result is what
`oracle/model.py`, `nerf-tf2_b200/model.py` and `include/nerfb200.h` hard-code. (Restated, not executed against
Keras: "parity unpinned" stays in force for this claim; what the test buys is that the claim is derived, step by
step, instead of remembered.)
"""
import collections

from oracle import model as om
import nerf_tf2_b200 as nb


class Layer:
    def __init__(self, name, has_weights=False):
        self.name, self.has_weights, self.inbound = name, has_weights, None   # one inbound node per layer in this model

    def __call__(self, *tensors):
        self.inbound = Node(self, tensors)
        return Tensor(self)


class Node:
    def __init__(self, layer, inputs):
        self.layer, self.inputs = layer, inputs
        self.parent_nodes = [t.layer.inbound for t in inputs]


class Tensor:
    def __init__(self, layer):
        self.layer = layer


def reference_graph(model_name):
    """The calls of get_coarse_or_fine_model (core/model.py:352-392), in the order they are made."""
    created = []

    def L(name, w=False):
        layer = Layer(f"{model_name}/{name}", w)
        created.append(layer)
        return layer

    xyz = L("xyz")()                                   # Input          :354
    rays_d = L("rays_d")()                             # Input          :355
    enc_xyz = L("enc_xyz")(xyz)                        # :358
    enc_rays_d = L("enc_rays_d")(rays_d)               # :361
    value = enc_xyz
    for i in range(8):                                 # :365-372
        value = L(f"dense_{i}", True)(value)
        if i == 4:
            value = L("concat_1")(value, enc_xyz)
    sigma = L("sigma", True)(value)                    # :375
    bottleneck = L("dense_8", True)(value)             # :378
    value = L("concat_2")(bottleneck, enc_rays_d)      # :381
    value = L("dense_9", True)(value)                  # :384
    rgb = L("rgb", True)(value)                        # :387
    return [xyz, rays_d], [rgb, sigma], created        # outputs = [rgb, sigma]   :390


def keras_layer_order(outputs):
    """Functional._map_graph_network restricted to what fixes `model.layers`."""
    # _build_map: depth-first from each output in turn; a layer's traversal index is assigned when it is FIRST met
    # (before its inputs are followed), a node is appended when all of its inputs are finished (post-order)
    finished, layer_indices, post_order = set(), {}, []

    def build_map(tensor):
        node = tensor.layer.inbound
        if node in finished:
            return
        layer_indices.setdefault(node.layer, len(layer_indices))
        for t in node.inputs:
            build_map(t)
        finished.add(node)
        post_order.append(node)

    for out in outputs:
        build_map(out)
    # depths: walk the nodes from the outputs back (reversed post-order puts every consumer before its producers);
    # depth of a node = longest path to an output
    nodes_depths, layers_depths = {}, {}
    for node in reversed(post_order):
        depth = max(nodes_depths.setdefault(node, 0), layers_depths.get(node.layer, 0))
        layers_depths[node.layer] = depth
        nodes_depths[node] = depth
        for parent in node.parent_nodes:
            nodes_depths[parent] = max(depth + 1, nodes_depths.get(parent, 0))
    by_depth = collections.defaultdict(list)
    for layer, depth in layers_depths.items():
        by_depth[depth].append(layer)
    layers = []
    for depth in sorted(by_depth, reverse=True):
        layers.extend(sorted(by_depth[depth], key=lambda l: layer_indices[l]))
    return layers, layers_depths


def test_variable_order_follows_from_keras_depth_sort():
    for model_name in ("coarse", "fine"):
        _, outputs, created = reference_graph(model_name)
        layers, depths = keras_layer_order(outputs)
        assert len(layers) == len(created) == 18
        weighted = [l.name.split("/", 1)[1] for l in layers if l.has_weights]
        assert weighted == om.LAYER_NAMES == nb.model.LAYER_NAMES == [f"dense_{i}" for i in range(10)] + ["rgb", "sigma"]
        # it differs from the creation order, which is why it matters
        assert [l.name.split("/", 1)[1] for l in created if l.has_weights] == \
            [f"dense_{i}" for i in range(8)] + ["sigma", "dense_8", "dense_9", "rgb"]
        by_name = {l.name.split("/", 1)[1]: d for l, d in depths.items()}
        # dense_7 feeds both sigma (depth 0) and dense_8 (depth 3): its depth is the LONGER way to an output
        assert by_name["dense_7"] == 4 and by_name["dense_8"] == 3 and by_name["dense_9"] == 1
        assert by_name["rgb"] == by_name["sigma"] == 0            # the tie that the traversal order breaks: rgb is output 0
        assert by_name["dense_0"] == 12 and by_name["enc_xyz"] == 13
        # the oracle's and the product's variable names follow the same order, kernel before bias
        names = om.variable_names(model_name)
        assert names == nb.model.variable_names(model_name)
        assert names[:2] == [f"{model_name}/dense_0/kernel", f"{model_name}/dense_0/bias"] and names[-2:] == \
            [f"{model_name}/sigma/kernel", f"{model_name}/sigma/bias"]


def test_swapping_the_outputs_would_swap_the_heads():
    """The order of the two heads is decided by `outputs = [rgb, sigma]` (core/model.py:390) alone."""
    _, outputs, _ = reference_graph("coarse")
    layers, _ = keras_layer_order(list(reversed(outputs)))
    weighted = [l.name.split("/", 1)[1] for l in layers if l.has_weights]
    assert weighted[-2:] == ["sigma", "rgb"] and weighted[:10] == [f"dense_{i}" for i in range(10)]
