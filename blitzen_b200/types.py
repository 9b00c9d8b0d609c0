"""POD layouts of the reference's frontend/backend contract, as numpy structured dtypes.

Every dtype mirrors one struct the reference memcpy's to the GPU unchanged
(/root/reference/src/Renderer/Resources/renderingResourcesTypes.h and Game/blitCamera.h):

  RenderObject        renderingResourcesTypes.h:155-159   8 B
  MeshTransform       renderingResourcesTypes.h:124-129  32 B
  PrimitiveSurface    renderingResourcesTypes.h:104-116  32 B
  LodData             renderingResourcesTypes.h:73-96    32 B
  Cluster             renderingResourcesTypes.h:27-47    32 B
  LodInstanceCounter  renderingResourcesTypes.h:98-102    8 B
  CameraViewData      Game/blitCamera.h:38-64           256 B (184 used)
  IndirectDraw (VK)   VulkanShaderHeaders/ShaderBuffers.glsl:73-84 == BlitzenVulkan/vulkanData.h:429-433  24 B
  DrawCmd (DX12)      HlslShaders/Headers/cullBuffers.hlsl:1-15 == BlitzenDX12/dx12Data.h:284-295         32 B
  ClusterDispatchData VulkanShaderHeaders/CullingShaderData.glsl:118-123 == vulkanData.h:442-447          12 B
"""
import numpy as np

RenderObject = np.dtype([("transformId", "<u4"), ("surfaceId", "<u4")])
MeshTransform = np.dtype([("pos", "<f4", (3,)), ("scale", "<f4"), ("orientation", "<f4", (4,))])
PrimitiveSurface = np.dtype([("center", "<f4", (3,)), ("radius", "<f4"), ("materialId", "<u4"),
                             ("lodOffset", "<u4"), ("lodCount", "<u4"), ("vertexOffset", "<u4")])
LodData = np.dtype([("indexCount", "<u4"), ("firstIndex", "<u4"), ("clusterOffset", "<u4"), ("clusterCount", "<u4"),
                    ("error", "<f4"), ("padding0", "<u4"), ("padding1", "<u4"), ("padding2", "<u4")])
Cluster = np.dtype([("center", "<f4", (3,)), ("radius", "<f4"),
                    ("coneAxisX", "i1"), ("coneAxisY", "i1"), ("coneAxisZ", "i1"), ("coneCutoff", "i1"),
                    ("dataOffset", "<u4"), ("vertexCount", "u1"), ("triangleCount", "u1"),
                    ("padding0", "u1"), ("padding1", "u1"), ("_tail", "<u4", (1,))])
LodInstanceCounter = np.dtype([("instanceOffset", "<u4"), ("instanceCount", "<u4")])
CameraViewData = np.dtype([("viewMatrix", "<f4", (16,)), ("projectionViewMatrix", "<f4", (16,)), ("position", "<f4", (3,)),
                           ("frustumRight", "<f4"), ("frustumLeft", "<f4"), ("frustumTop", "<f4"), ("frustumBottom", "<f4"),
                           ("proj0", "<f4"), ("proj5", "<f4"), ("zNear", "<f4"), ("zFar", "<f4"),
                           ("pyramidWidth", "<f4"), ("pyramidHeight", "<f4"), ("lodTarget", "<f4"),
                           ("_pad", "u1", (72,))])
IndirectDrawVK = np.dtype([("objectId", "<u4"), ("indexCount", "<u4"), ("instanceCount", "<u4"),
                           ("firstIndex", "<u4"), ("vertexOffset", "<u4"), ("firstInstance", "<u4")])
DrawCmdDX = np.dtype([("objId", "<u4"), ("indexCount", "<u4"), ("instCount", "<u4"), ("indexOffset", "<u4"),
                      ("vertOffset", "<i4"), ("insOffset", "<u4"), ("padding0", "<u4"), ("padding1", "<u4")])
ClusterDispatchData = np.dtype([("objectId", "<u4"), ("lodIndex", "<u4"), ("clusterId", "<u4")])

assert RenderObject.itemsize == 8 and MeshTransform.itemsize == 32 and PrimitiveSurface.itemsize == 32
assert LodData.itemsize == 32 and Cluster.itemsize == 32 and LodInstanceCounter.itemsize == 8
assert CameraViewData.itemsize == 256 and IndirectDrawVK.itemsize == 24 and DrawCmdDX.itemsize == 32
assert ClusterDispatchData.itemsize == 12
