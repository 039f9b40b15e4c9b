#include "../../include/b200rc_host.h"
#include <cstdio>
int main(){ b200host_svrc* p; printf("create %d\n", b200host_svrc_create(0,&p));
 printf("%d\n", b200host_svrc_set_property(p,"compositing","dvr"));
 printf("%d\n", b200host_svrc_set_property(p,"classification","pre-integrated-cpu"));
 b200rc_volume_desc d; int32_t dims[3]={40,40,40}; b200host_volume_desc(dims,1,"uint16",0,0,0,&d);
 static unsigned short v[64000]; printf("vol %d\n", b200host_svrc_set_volume(p,v,&d));
 b200host_tf_key k[2]={{0.f,{0,0,0,0},{0,0,0,0},0},{1.f,{255,255,255,255},{255,255,255,255},0}};
 printf("tf %d\n", b200host_svrc_set_transfer_function(p,0,k,2,0,65535,0,1.f));
 float e[3]={1,1,3},f[3]={0,0,0},u[3]={0,1,0}; printf("cam %d\n", b200host_svrc_set_camera(p,e,f,u,45,0.1,50));
 void* outs[3]={0,0,0}; printf("ready %d proc %d\n", b200host_svrc_is_ready(p), b200host_svrc_process(p,64,64,outs,0));
 b200host_svrc_destroy(p); return 0; }
