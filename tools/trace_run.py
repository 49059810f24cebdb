import torch, fourier_b200 as fb
p=fb.create_fft_f32(1<<20)
x=torch.empty((256,1<<20),dtype=torch.complex64,device='cuda'); fb.fill_input(x); y=torch.empty_like(x)
p.transform(x,y,fb.Transform.Fft); torch.cuda.synchronize()
