/* ocb_kdata.cpp -- the <orcdchomp><spheres> block of a robot / kinbody XML file: the sphere table
 * mod::create walks (src/orcdchomp_mod.cpp:2180-2210), read the way the reference's XML reader does
 * (src/orcdchomp_kdata.cpp:65-98: <sphere link="..." pos="x y z" radius="r"/> inside <spheres>;
 * any other attribute is an error, a <sphere> outside <spheres> is ignored with an error message,
 * nested <spheres> is an error).  The reference hooks into OpenRAVE's SAX reader; here the same
 * element / attribute handling runs over a small tag scanner (comments, processing instructions and
 * everything outside <orcdchomp> are skipped).  Host code only. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/orcdchomp_b200_module.h"

namespace
{
struct Attr { std::string name, value; };

/* next tag at or after *pos: name ("/name" for a closing tag), attributes, self-closing flag */
bool next_tag(const std::string &s, size_t *pos, std::string &name, std::vector<Attr> &atts, bool &self_closing)
{
   for (;;)
   {
      size_t lt = s.find('<', *pos);
      if (lt == std::string::npos) return false;
      if (s.compare(lt, 4, "<!--") == 0)
      {
         size_t e = s.find("-->", lt + 4);
         if (e == std::string::npos) return false;
         *pos = e + 3;
         continue;
      }
      if (lt + 1 < s.size() && (s[lt + 1] == '?' || s[lt + 1] == '!'))
      {
         size_t e = s.find('>', lt);
         if (e == std::string::npos) return false;
         *pos = e + 1;
         continue;
      }
      size_t gt = lt + 1;
      char quote = 0;
      for (; gt < s.size(); gt++)
      {
         if (quote) { if (s[gt] == quote) quote = 0; }
         else if (s[gt] == '"' || s[gt] == '\'') quote = s[gt];
         else if (s[gt] == '>') break;
      }
      if (gt >= s.size()) return false;
      std::string body = s.substr(lt + 1, gt - lt - 1);
      *pos = gt + 1;
      self_closing = !body.empty() && body.back() == '/';
      if (self_closing) body.pop_back();
      size_t k = 0;
      while (k < body.size() && !isspace((unsigned char) body[k])) k++;
      name = body.substr(0, k);
      atts.clear();
      while (k < body.size())
      {
         while (k < body.size() && isspace((unsigned char) body[k])) k++;
         size_t e = k;
         while (e < body.size() && body[e] != '=' && !isspace((unsigned char) body[e])) e++;
         if (e == k) break;
         Attr a;
         a.name = body.substr(k, e - k);
         k = e;
         while (k < body.size() && isspace((unsigned char) body[k])) k++;
         if (k < body.size() && body[k] == '=')
         {
            k++;
            while (k < body.size() && isspace((unsigned char) body[k])) k++;
            if (k < body.size() && (body[k] == '"' || body[k] == '\''))
            {
               const char q = body[k++];
               size_t e2 = body.find(q, k);
               if (e2 == std::string::npos) e2 = body.size();
               a.value = body.substr(k, e2 - k);
               k = e2 + 1;
            }
         }
         atts.push_back(a);
      }
      return true;
   }
}
} /* namespace */

/* Parses every <sphere> of the <orcdchomp><spheres> block(s) of `xml`, in document order.
 * link_names: [cap][64], pos: [cap][3], radius: [cap].  *n_out = number found (may exceed cap: only the
 * first cap are stored).  Returns 0, or -2 with the reference's message in ocb_module_last_error(). */
extern "C" int ocb_kdata_parse_spheres(const char *xml, int cap, char *link_names, double *pos, double *radius,
                                       int *n_out, char *err, size_t err_cap)
{
   if (!xml || !n_out || cap < 0) return -2;
   auto fail = [&](const std::string &m)
   {
      if (err && err_cap) snprintf(err, err_cap, "%s", m.c_str());
      return -2;
   };
   const std::string s(xml);
   size_t p = 0;
   std::string name;
   std::vector<Attr> atts;
   bool selfc = false, inside_kdata = false, inside_spheres = false;
   int n = 0;
   while (next_tag(s, &p, name, atts, selfc))
   {
      if (name == "orcdchomp") { inside_kdata = !selfc; continue; }
      if (name == "/orcdchomp") { inside_kdata = false; inside_spheres = false; continue; }
      if (!inside_kdata) continue;
      if (name == "spheres")
      {
         if (inside_spheres) return fail("you can't have <spheres> inside <spheres>!"); /* kdata.cpp:69 */
         inside_spheres = !selfc;
      }
      else if (name == "/spheres")
      {
         if (!inside_spheres) return fail("you can't have </spheres> without matching <spheres>!"); /* kdata.cpp:110 */
         inside_spheres = false;
      }
      else if (name == "sphere")
      {
         if (!inside_spheres) continue; /* "you can't have <sphere> not inside <spheres>!": ignored (kdata.cpp:77) */
         char link[64] = "";
         double c[3] = {0.0, 0.0, 0.0}, r = 0.0;
         for (const Attr &a : atts)
         {
            if (a.name == "link") snprintf(link, sizeof(link), "%s", a.value.c_str());
            else if (a.name == "radius") r = strtod(a.value.c_str(), 0);
            else if (a.name == "pos") sscanf(a.value.c_str(), "%lf %lf %lf", &c[0], &c[1], &c[2]);
            else return fail("unknown attribute " + a.name + "=" + a.value + "!"); /* kdata.cpp:89 */
         }
         if (n < cap)
         {
            if (link_names) memcpy(link_names + (size_t) n * 64, link, 64);
            if (pos) { pos[3 * n] = c[0]; pos[3 * n + 1] = c[1]; pos[3 * n + 2] = c[2]; }
            if (radius) radius[n] = r;
         }
         n++;
      }
   }
   *n_out = n;
   return 0;
}
