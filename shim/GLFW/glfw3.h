/* shim/GLFW/glfw3.h -- the few GLFW declarations the reference's loader files and camera.c need (boundary B1).
 * src/vulkan_basics.h:17 includes <GLFW/glfw3.h> with GLFW_INCLUDE_VULKAN; src/camera.c:106-131 polls keys and the mouse.
 * No window system is involved: the input functions report "nothing pressed", time stands still. */
#ifndef VKR_SHIM_GLFW3_H
#define VKR_SHIM_GLFW3_H
#include <vulkan/vulkan.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct GLFWwindow GLFWwindow;
typedef void (*GLFWvkproc)(void);
#define GLFW_RELEASE 0
#define GLFW_PRESS 1
#define GLFW_MOUSE_BUTTON_1 0
#define GLFW_MOUSE_BUTTON_2 1
#define GLFW_KEY_A 65
#define GLFW_KEY_D 68
#define GLFW_KEY_E 69
#define GLFW_KEY_Q 81
#define GLFW_KEY_S 83
#define GLFW_KEY_W 87
#define GLFW_KEY_LEFT_SHIFT 340
#define GLFW_KEY_LEFT_CONTROL 341
GLFWvkproc glfwGetInstanceProcAddress(VkInstance instance, const char* procname);
int glfwGetKey(GLFWwindow* window, int key);
int glfwGetMouseButton(GLFWwindow* window, int button);
void glfwGetCursorPos(GLFWwindow* window, double* xpos, double* ypos);
double glfwGetTime(void);
#ifdef __cplusplus
}
#endif
#endif
