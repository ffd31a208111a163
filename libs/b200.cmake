# libs/b200.cmake -- included when ENABLE_B200 is ON (pattern:
# libs/cusparse.cmake:1-5 of the reference: enable CUDA, find the toolkit, link
# the runtime).  Builds libb200 from the hand-written sm_100a kernels; NCCL is
# bound at run time with dlopen, so only its header is needed here.
enable_language(CUDA)
find_package(CUDAToolkit 12.8 REQUIRED)
find_package(Threads REQUIRED)

set(B200_CSRC ${CMAKE_CURRENT_LIST_DIR}/../lsbench_b200/csrc)
add_library(b200 SHARED
  ${B200_CSRC}/ctx.cu ${B200_CSRC}/convert.cu ${B200_CSRC}/spmv.cu
  ${B200_CSRC}/pcg.cu ${B200_CSRC}/generate.cu ${B200_CSRC}/dist.cu
  ${B200_CSRC}/small.cu ${B200_CSRC}/ingest.cu)
target_include_directories(b200 PUBLIC ${CMAKE_CURRENT_LIST_DIR}/../include
  PRIVATE ${B200_CSRC})
# sm_100a only: no multi-arch dispatch
set_target_properties(b200 PROPERTIES CUDA_ARCHITECTURES "100a"
  CUDA_STANDARD 17 POSITION_INDEPENDENT_CODE ON PUBLIC_HEADER include/b200.h)
target_compile_options(b200 PRIVATE $<$<COMPILE_LANGUAGE:CUDA>:-lineinfo -O3>)
target_link_libraries(b200 PRIVATE CUDA::cudart_static ${CMAKE_DL_LIBS} Threads::Threads)

target_link_libraries(lsbench PRIVATE b200 Threads::Threads)
install(TARGETS b200 LIBRARY DESTINATION lib PUBLIC_HEADER DESTINATION include)
