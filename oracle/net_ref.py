"""L0 oracle: the four reference networks in plain torch fp32 on CPU, from the exported blob.

Follows the TorchScript graphs of net/{Backbone,PointHeatmap,EdgeHeatmap,Descriptor}.pt as the
reference runs them in feature/src/PPGExtractor.cpp:149-156 (inference), :161-162 (junction
softmax + pixel_shuffle) and :242 (heat softmax).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import numpy as np
import torch
import torch.nn.functional as F

from ppg_slam_b200.weights_io import load_blob


class NetRef:
    def __init__(self, blob_path=None, threads=None):
        w = load_blob(blob_path) if blob_path else load_blob()
        self.w = {k: torch.from_numpy(np.array(v)) for k, v in w.items()}
        if threads:
            torch.set_num_threads(threads)

    def _conv(self, x, name, pad=1):
        return F.conv2d(x, self.w[name + ".weight"], self.w[name + ".bias"], padding=pad)

    @torch.no_grad()
    def backbone(self, x):
        # Backbone.pt: SuperpointBackbone.forward
        x = F.relu(self._conv(x, "backbone.conv1a"))
        x = F.relu(self._conv(x, "backbone.conv1b"))
        x = F.max_pool2d(x, 2, 2)
        x = F.relu(self._conv(x, "backbone.conv2a"))
        x = F.relu(self._conv(x, "backbone.conv2b"))
        x = F.max_pool2d(x, 2, 2)
        x = F.relu(self._conv(x, "backbone.conv3a"))
        x = F.relu(self._conv(x, "backbone.conv3b"))
        x = F.max_pool2d(x, 2, 2)
        x = F.relu(self._conv(x, "backbone.conv4a"))
        x = F.relu(self._conv(x, "backbone.conv4b"))
        return x

    @torch.no_grad()
    def junction(self, f):
        return self._conv(F.relu(self._conv(f, "junction.convPa")), "junction.convPb", pad=0)

    @torch.no_grad()
    def descriptor(self, f):
        return self._conv(F.relu(self._conv(f, "descriptor.convDa")), "descriptor.convDb", pad=0)

    @torch.no_grad()
    def edge(self, f):
        x = f
        for i in range(3):
            p = "edge.conv_block_lst.%d" % i
            x = self._conv(x, p + ".0")
            x = F.batch_norm(x, self.w[p + ".1.running_mean"], self.w[p + ".1.running_var"],
                             self.w[p + ".1.weight"], self.w[p + ".1.bias"], False, 0.1, 1e-5)
            x = F.pixel_shuffle(F.relu(x), 2)
        return self._conv(x, "edge.conv_block_lst.3", pad=0)

    @torch.no_grad()
    def forward_u8(self, gray_u8):
        """gray (H,W) u8 -> dict of the dense maps the post-processing consumes.

        prob  (H,W) f32   PPGExtractor.cpp:161-162  softmax(junctions,1)[:, :64] -> pixel_shuffle(8)
        heat  (H,W) f32   PPGExtractor.cpp:242      softmax(heatmap,1)[:,1]
        desc  (256,Hc,Wc) PPGExtractor.cpp:155      raw dense descriptors
        """
        x = torch.from_numpy(np.ascontiguousarray(gray_u8))[None, None].to(torch.float32) / 255.0
        f = self.backbone(x)
        j = self.junction(f)
        h = self.edge(f)
        d = self.descriptor(f)
        prob = F.pixel_shuffle(torch.softmax(j, 1).narrow(1, 0, 64), 8)[0, 0]
        heat = torch.softmax(h, 1).select(1, 1)[0]
        return dict(feature=f[0].numpy(), junction_logits=j[0].numpy(), heat_logits=h[0].numpy(),
                    prob=prob.contiguous().numpy(), heat=heat.contiguous().numpy(),
                    desc=d[0].contiguous().numpy())
